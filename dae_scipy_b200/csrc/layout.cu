// layout.cu -- tiled transposes between the row-per-system layout of a solve_ivp-style caller
// ([B][n] states, [B][n][T] results) and the structure-of-arrays layout of the kernels
// ([n][B], [T][n][B]).  HBM-bound byte shuffles: 32x32 tiles through padded shared memory so that
// both the read and the write side are coalesced.
#include "kernels.cuh"

namespace brdae {

template <class T>
__global__ void permute_kernel(const T *__restrict__ src, T *__restrict__ dst, int64_t R, int64_t C, int64_t src_rs,
                               int64_t dst_rs, int64_t src_zs, int64_t dst_zs) {
    __shared__ T tile[32][33];
    const int64_t z = blockIdx.z;
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    src += z * src_zs;
    dst += z * dst_zs;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = r0 + j, c = c0 + threadIdx.x;
        if (r < R && c < C) tile[j][threadIdx.x] = src[r * src_rs + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t c = c0 + j, r = r0 + threadIdx.x;
        if (r < R && c < C) dst[c * dst_rs + r] = tile[threadIdx.x][j];
    }
}

template <class T>
static void permute(const T *src, T *dst, int64_t R, int64_t C, int64_t src_rs, int64_t dst_rs, int Z, int64_t src_zs,
                    int64_t dst_zs, cudaStream_t s) {
    if (R <= 0 || C <= 0 || Z <= 0) return;
    // the long axis goes on grid.x (2^31 limit); grid.y is limited to 65535
    const int64_t gx = (C + 31) / 32, gy = (R + 31) / 32;
    if (gy <= 65535) {
        dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)Z), block(32, 8);
        permute_kernel<T><<<grid, block, 0, s>>>(src, dst, R, C, src_rs, dst_rs, src_zs, dst_zs);
    } else {
        // swap the roles: transposing (C x R) of the destination view is the same permutation
        for (int64_t r0 = 0; r0 < R; r0 += 65535LL * 32) {
            const int64_t rr = (R - r0 < 65535LL * 32) ? R - r0 : 65535LL * 32;
            dim3 grid((unsigned)gx, (unsigned)((rr + 31) / 32), (unsigned)Z), block(32, 8);
            permute_kernel<T><<<grid, block, 0, s>>>(src + r0 * src_rs, dst + r0, rr, C, src_rs, dst_rs, src_zs, dst_zs);
        }
    }
}

// how many entries of v[0..n) are NaN or +-inf, added to *count (scipy's check of y0, base.py:14-17, for the host-buffer entry:
// the state is on the device anyway, and summing 40 MB on one host core cost 4 % of an end-to-end step)
__global__ void count_nonfinite_kernel(const double *__restrict__ v, int64_t n, unsigned long long *count) {
    unsigned long long bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        bad += !(fabs(v[i]) <= 1.7976931348623157e308);
    bad = __reduce_add_sync(0xffffffffu, (unsigned)bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(count, bad);
}
void count_nonfinite(const double *v, int64_t n, unsigned long long *count, cudaStream_t s) {
    if (n <= 0) return;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    count_nonfinite_kernel<<<(unsigned)blocks, 256, 0, s>>>(v, n, count);
}

void permute_f64(const double *src, double *dst, int64_t R, int64_t C, int64_t src_rs, int64_t dst_rs, int Z,
                 int64_t src_zs, int64_t dst_zs, cudaStream_t s) {
    permute<double>(src, dst, R, C, src_rs, dst_rs, Z, src_zs, dst_zs, s);
}
void permute_i32(const int32_t *src, int32_t *dst, int64_t R, int64_t C, int64_t src_rs, int64_t dst_rs, int Z,
                 int64_t src_zs, int64_t dst_zs, cudaStream_t s) {
    permute<int32_t>(src, dst, R, C, src_rs, dst_rs, Z, src_zs, dst_zs, s);
}

}  // namespace brdae
