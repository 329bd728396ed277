// k_march.cuh -- placeholder until the x-marching register-pipelined kernel lands.
#pragma once
#include "k_naive.cuh"
namespace phb {
template <class T> inline bool march_supported(int, int, int, int) { return false; }
template <class T> inline const char *march_name() { return "march"; }
template <class A, class M>
inline int launch_march(const StepArgs<typename A::T> &, const M &, cudaStream_t) { return 0; }
}  // namespace phb
