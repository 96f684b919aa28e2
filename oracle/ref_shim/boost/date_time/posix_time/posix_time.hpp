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
// Every reading of the clock is also kept (up to 64 per run) so that the driver can report the reference's OWN
// phases with sub-second resolution: runCoGAPSAlgorithm reads it at the start of loading, at its end
// (GapsRunner.cpp:400,414), when the sampler loop starts (:450) and when it ends (:473), plus once per status
// line; GapsResult::totalRunningTime itself is truncated to whole seconds.
struct clock_marks
{
    std::chrono::steady_clock::time_point t[64];
    unsigned n;
};
inline clock_marks &marks() { static clock_marks m = clock_marks(); return m; }
struct microsec_clock
{
    static ptime local_time()
    {
        ptime p;
        p.t = std::chrono::steady_clock::now();
        clock_marks &m = marks();
        if (m.n < 64) { m.t[m.n] = p.t; }
        ++m.n;
        m.t[63] = p.t; // the latest reading is always kept in the last slot
        return p;
    }
};
}} // namespace boost::posix_time
#endif
