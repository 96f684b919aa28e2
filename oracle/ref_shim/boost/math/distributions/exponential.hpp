// Included by math/Math.cpp:4 but never used there.  Test infrastructure only.
#ifndef COGAPS_B200_SHIM_EXPONENTIAL_HPP
#define COGAPS_B200_SHIM_EXPONENTIAL_HPP
#endif
