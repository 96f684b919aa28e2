// Minimal stand-in for boost::math::gamma_distribution (math/Math.cpp:43-63).  The reference
// only ever asks for shape 2 (math/Random.cpp:288-294); other shapes go through a series/
// continued-fraction regularised incomplete gamma.  Test infrastructure only.
#ifndef COGAPS_B200_SHIM_GAMMA_HPP
#define COGAPS_B200_SHIM_GAMMA_HPP
#include <cmath>
namespace boost { namespace math {
template <class RealType = double>
struct gamma_distribution
{
    RealType k, theta;
    gamma_distribution(RealType shape, RealType scale = 1) : k(shape), theta(scale) {}
};
namespace shim_detail {
inline double reg_lower_gamma(double a, double x)
{
    if (x <= 0.0) { return 0.0; }
    if (a == 2.0) { return 1.0 - (1.0 + x) * std::exp(-x); }
    if (x < a + 1.0)
    {
        double ap = a, sum = 1.0 / a, del = sum;
        for (int n = 0; n < 1000; ++n)
        {
            ap += 1.0; del *= x / ap; sum += del;
            if (std::fabs(del) < std::fabs(sum) * 1e-17) { break; }
        }
        return sum * std::exp(-x + a * std::log(x) - ::lgamma(a));
    }
    double b = x + 1.0 - a, c = 1.0 / 1e-300, dd = 1.0 / b, h = dd;
    for (int i = 1; i < 1000; ++i)
    {
        double an = -i * (i - a);
        b += 2.0;
        dd = an * dd + b; if (std::fabs(dd) < 1e-300) { dd = 1e-300; }
        c = b + an / c;   if (std::fabs(c) < 1e-300) { c = 1e-300; }
        dd = 1.0 / dd;
        double del = dd * c; h *= del;
        if (std::fabs(del - 1.0) < 1e-17) { break; }
    }
    return 1.0 - std::exp(-x + a * std::log(x) - ::lgamma(a)) * h;
}
} // namespace shim_detail
template <class R, class X> inline R pdf(const gamma_distribution<R> &d, X xin)
{
    R x = static_cast<R>(xin);
    if (x < 0) { return 0; }
    return std::exp((d.k - 1) * std::log(x / d.theta) - x / d.theta - ::lgamma(d.k)) / d.theta;
}
template <class R, class X> inline R cdf(const gamma_distribution<R> &d, X xin)
{
    R x = static_cast<R>(xin);
    return shim_detail::reg_lower_gamma(d.k, x / d.theta);
}
template <class R, class X> inline R quantile(const gamma_distribution<R> &d, X xin)
{
    R p = static_cast<R>(xin);
    R lo = 0.0, hi = 1.0;
    while (shim_detail::reg_lower_gamma(d.k, hi) < p && hi < 1e300) { hi *= 2.0; }
    for (int it = 0; it < 400; ++it)
    {
        R mid = 0.5 * (lo + hi);
        if (mid == lo || mid == hi) { break; }
        if (shim_detail::reg_lower_gamma(d.k, mid) < p) { lo = mid; } else { hi = mid; }
    }
    return d.theta * (0.5 * (lo + hi));
}
}} // namespace boost::math
#endif
