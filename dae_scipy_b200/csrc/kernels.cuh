// kernels.cuh -- host-side interface between the C-ABI translation unit and the kernel
// translation units (one per problem family so that they compile in parallel).
#pragma once
#include <cuda_runtime.h>

#include "radau_common.cuh"

namespace brdae {

struct LaunchGeom { int grid, block, smem; size_t scratch_bytes; };   // scratch_bytes: K1 cold-state slice of the workspace

// geometry that launch_solver would use (no launch)
int launch_geometry(int problem, bool dense_mass, bool jac_fd, int n, int64_t B, int sms, LaunchGeom *g);
// launch the kernel matching (problem, mass kind) on `stream`
int launch_solver(int problem, bool dense_mass, const SolveArgs &a, const brdae_batch *b, int sms, cudaStream_t stream);

// per-problem launchers (k1_*.cu); geom_only: fill *g and return without launching
int k1_pendulum(int index, bool dense_mass, const SolveArgs &a, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only);
int k1_robertson(bool dense_mass, const SolveArgs &a, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only);
int k1_transistor(bool dense_mass, const SolveArgs &a, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only);
int k1_extra(int problem, bool dense_mass, const SolveArgs &a, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only);
// the user-supplied functor of a plugin build (k1_user.cu); BRDAE_ERR_UNKNOWN_PROBLEM in the stock library
int k1_user(bool dense_mass, const SolveArgs &a, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only);
// CTA-per-system banded kernel for the n-pendulum chain (k2_npendulum.cu)
int k2_npendulum(bool dense_mass, const SolveArgs &a, const brdae_batch *b, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only);
size_t k2_scratch_bytes(int n, int sms, int64_t B, bool fd);
// the same kernel for a user band problem with more than 8 unknowns (k2_user.cu; a stub in the stock library)
void count_nonfinite(const double *v, int64_t n, unsigned long long *count, cudaStream_t s);
int k2_user(bool dense_mass, const SolveArgs &a, const brdae_batch *b, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only);
size_t k2_user_scratch_bytes(int n, int sms, int64_t B, bool fd);
// lane-per-rope block-elimination kernel for the chain (k4_chain.cu, chain_k4.cuh)
int k4_chain(bool dense_mass, const SolveArgs &a, const brdae_batch *b, int sms, cudaStream_t s, LaunchGeom *g, bool geom_only);
size_t k4_scratch_bytes(int n, int64_t B, int sms);
// which chain kernel a batch of B ropes with n unknowns uses (BRDAE_NPENDULUM_KERNEL=k2 selects the band kernel)
bool npendulum_uses_k4(int n, int64_t B);

// layout.cu: dst[z*dst_zs + c*dst_rs + r] = src[z*src_zs + r*src_rs + c] for r < R, c < C, z < Z
// (rows <-> structure-of-arrays permutations of the host-buffer entry points)
void permute_f64(const double *src, double *dst, int64_t R, int64_t C, int64_t src_rs, int64_t dst_rs, int Z,
                 int64_t src_zs, int64_t dst_zs, cudaStream_t s);
void permute_i32(const int32_t *src, int32_t *dst, int64_t R, int64_t C, int64_t src_rs, int64_t dst_rs, int Z,
                 int64_t src_zs, int64_t dst_zs, cudaStream_t s);

int fp64_peak_probe(double ms_target, double *flops, cudaStream_t s, int sms);

}  // namespace brdae
