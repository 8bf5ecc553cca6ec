// oracle/shim/rrt_seeded_random.h — TEST INFRASTRUCTURE, force-included (-include) when the reference's rrt_star.cpp is compiled.
// RRTStar::getRandomNode (rrt_star.cpp:104-116) constructs a std::random_device and a std::mt19937_64 seeded with ONE 32-bit draw from
// it for EVERY sample, so the reference's sample i is a pure function of the 32-bit value random_device returned for it.  To make the
// reference's own code reproducible, the token `random_device` is redirected (after <random> has been included for real) to a class
// that returns seed32(query_seed, i) — the same counter-based stream the CUDA kernel and the restatement use (uavmp.h:
// uavmp_rrt_sample_seed) — and counts the samples drawn, which is also the deterministic clock the driver installs in ros::Time
// (the reference stops on wall-clock time, rrt_star.cpp:413-418).
#pragma once
#include <random>

extern "C" unsigned long long rrt_shim_query_seed;  // defined in oracle/rrt_ref_driver.cpp
extern "C" long long rrt_shim_samples;              // number of random_device objects constructed = samples drawn so far

static inline unsigned int rrt_shim_seed32(unsigned long long query_seed, long long i) {
  unsigned long long z = query_seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (unsigned int)(z >> 32);
}

namespace std {
class uavmp_seeded_random_device {
 public:
  typedef unsigned int result_type;
  uavmp_seeded_random_device() : i_(rrt_shim_samples++) {}
  result_type operator()() { return rrt_shim_seed32(rrt_shim_query_seed, i_); }
 private:
  long long i_;
};
}  // namespace std
#define random_device uavmp_seeded_random_device
