// Minimal stand-in for boost::math::normal_distribution + lgamma (math/Math.cpp:65-86).
// cdf through libm erfc in f64; quantile by f64 bisection run to a fixed point, so the
// float-rounded lookup tables are reproducible anywhere libm is.  Test infrastructure only.
#ifndef COGAPS_B200_SHIM_NORMAL_HPP
#define COGAPS_B200_SHIM_NORMAL_HPP
#include <cmath>
namespace boost { namespace math {
template <class RealType = double>
struct normal_distribution
{
    RealType m, s;
    normal_distribution(RealType mean = 0, RealType sd = 1) : m(mean), s(sd) {}
};
template <class R, class X> inline R pdf(const normal_distribution<R> &d, X xin)
{
    R x = static_cast<R>(xin);
    R z = (x - d.m) / d.s;
    return std::exp(-0.5 * z * z) / (d.s * 2.5066282746310005024157652848110452530069867406099);
}
template <class R, class X> inline R cdf(const normal_distribution<R> &d, X xin)
{
    R x = static_cast<R>(xin);
    R z = (x - d.m) / d.s;
    return 0.5 * std::erfc(-z / 1.4142135623730950488016887242096980785696718753769);
}
template <class R, class X> inline R quantile(const normal_distribution<R> &d, X xin)
{
    R p = static_cast<R>(xin);
    normal_distribution<R> unit(0, 1);
    R lo = -40.0, hi = 40.0;
    for (int it = 0; it < 400; ++it)
    {
        R mid = 0.5 * (lo + hi);
        if (mid == lo || mid == hi) { break; }
        if (cdf(unit, mid) < p) { lo = mid; } else { hi = mid; }
    }
    return d.m + d.s * (0.5 * (lo + hi));
}
inline double lgamma(double x) { return ::lgamma(x); }
}} // namespace boost::math
#endif
