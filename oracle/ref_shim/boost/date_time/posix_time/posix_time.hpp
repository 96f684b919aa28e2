// Minimal stand-in for the boost::posix_time wall clock used at GapsRunner.cpp:25-31.
// Test infrastructure only.
#ifndef COGAPS_B200_SHIM_POSIX_TIME_HPP
#define COGAPS_B200_SHIM_POSIX_TIME_HPP
#include <chrono>
namespace boost { namespace posix_time {
struct time_duration
{
    std::chrono::steady_clock::duration d;
    long total_seconds() const { return std::chrono::duration_cast<std::chrono::seconds>(d).count(); }
    long total_microseconds() const { return std::chrono::duration_cast<std::chrono::microseconds>(d).count(); }
};
struct ptime
{
    std::chrono::steady_clock::time_point t;
};
inline time_duration operator-(const ptime &a, const ptime &b) { time_duration r; r.d = a.t - b.t; return r; }
struct microsec_clock
{
    static ptime local_time() { ptime p; p.t = std::chrono::steady_clock::now(); return p; }
};
}} // namespace boost::posix_time
#endif
