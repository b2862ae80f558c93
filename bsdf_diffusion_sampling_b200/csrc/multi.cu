// One wavefront, several materials: the device-side plan.
//
// The reference's array scenes hold twelve `mybsdf` instances (rendering/matpreview/disney_bsdf_array0_envmap.xml:35-335);
// Mitsuba calls each instance with the lanes that hit it.  A wavefront renderer that keeps ONE batch with a material id
// per lane gets one launch instead: the rows are bucketed by material ON THE DEVICE (no host read of the bucket sizes)
// into "virtual tiles" of <= 128 rows of one material, and the sampler kernels (flow_tc.cu, flow_simt.cu) walk the tile
// table, switching the weight set in shared memory per tile.  Outputs are written straight to the wavefront row, and the
// Philox counter is the wavefront row as well, so a row's result does not depend on the bucketing at all: it equals the
// single-material call on the same row bit for bit.
//
// Plan layout in the caller's scratch buffer (32-bit words; bsdfdiff_multi_scratch_bytes):
//   [0, 256)           counts[m]          rows of material m; slot M counts the inactive rows (id outside [0, M))
//   [256, 512)         cursor[m]          scatter cursors (plan build only)
//   [512, 769)         seg_off[m]         first position of material m in the material-sorted order; [M] = first inactive
//                                         position, [M + 1] = n
//   [769]              n_tiles
//   [1024, 1280)       fix_count[m]       per-material fix-up list lengths (zeroed by every sample / pdf call)
//   [1280 ...)         tiles int4 x (n / 128 + M + 1)   {material, first position, rows, first position of the material}
//   then               perm u32 x n       wavefront row of every position
//   then               fix_list u32 x n   material m's flagged rows at [seg_off[m] ...)
//   then               x0 float2 x n      base samples kept for the fix-up pass when the caller supplies none
#include "common.cuh"
#include "multi.cuh"

namespace bsdfdiff {

constexpr int kPlanThreads = 256, kPlanItems = 8;      // rows per block = 2048

// rows per material (block-level shared-memory histogram, one global atomic per material per block)
__global__ void __launch_bounds__(kPlanThreads) multi_count_kernel(const int* __restrict__ mid, long long n, int M,
                                                                  unsigned int* __restrict__ counts) {
    __shared__ unsigned int h[kMaxMaterials + 1];
    for (int i = threadIdx.x; i <= M; i += kPlanThreads) h[i] = 0u;
    __syncthreads();
    const long long base = (long long)blockIdx.x * (kPlanThreads * kPlanItems);
#pragma unroll
    for (int it = 0; it < kPlanItems; ++it) {
        const long long i = base + it * kPlanThreads + threadIdx.x;
        if (i < n) {
            const int m = mid[i];
            atomicAdd(&h[(m >= 0 && m < M) ? m : M], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= M; i += kPlanThreads)
        if (h[i]) atomicAdd(&counts[i], h[i]);
}

// segment offsets + the virtual-tile table.  Every block redoes the (tiny) scan over the materials in shared memory and
// writes its own slice of the tile table, so the table of a 16 M-row wavefront (131 K tiles) is not one block's job.
__global__ void __launch_bounds__(kPlanThreads) multi_tiles_kernel(long long n, int M, const unsigned int* __restrict__ counts,
                                                                  unsigned int* __restrict__ seg_off,
                                                                  unsigned int* __restrict__ n_tiles, int4* __restrict__ tiles) {
    __shared__ unsigned int cnt[kMaxMaterials + 1], off[kMaxMaterials + 2], toff[kMaxMaterials + 1];
    for (int i = threadIdx.x; i < M; i += kPlanThreads) cnt[i] = counts[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int a = 0, t = 0;
        for (int m = 0; m < M; ++m) {
            off[m] = a; toff[m] = t;
            a += cnt[m];
            t += (cnt[m] + 127u) / 128u;
        }
        off[M] = a; off[M + 1] = (unsigned int)n; toff[M] = t;
        if (blockIdx.x == 0) *n_tiles = t;
    }
    __syncthreads();
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i <= M + 1; i += kPlanThreads) seg_off[i] = off[i];
    const unsigned int total = toff[M];
    for (unsigned int j = blockIdx.x * kPlanThreads + threadIdx.x; j < total; j += gridDim.x * kPlanThreads) {
        int lo = 0, hi = M - 1;                         // last material whose first tile is <= j (it has tiles: see below)
        while (lo < hi) {
            const int mid_ = (lo + hi + 1) >> 1;
            if (toff[mid_] <= j) lo = mid_; else hi = mid_ - 1;
        }
        const unsigned int first = off[lo] + (j - toff[lo]) * 128u;
        const unsigned int rows = min(128u, off[lo + 1] - first);
        tiles[j] = make_int4(lo, (int)first, (int)rows, (int)off[lo]);
    }
}

// perm[position] = wavefront row (block-aggregated: one global atomic per material per block reserves a range; rows of
// one material keep no particular order -- no result depends on it)
__global__ void __launch_bounds__(kPlanThreads) multi_scatter_kernel(const int* __restrict__ mid, long long n, int M,
                                                                    const unsigned int* __restrict__ seg_off,
                                                                    unsigned int* __restrict__ cursor,
                                                                    unsigned int* __restrict__ perm) {
    __shared__ unsigned int h[kMaxMaterials + 1], start[kMaxMaterials + 1];
    for (int i = threadIdx.x; i <= M; i += kPlanThreads) h[i] = 0u;
    __syncthreads();
    const long long base = (long long)blockIdx.x * (kPlanThreads * kPlanItems);
    int mm[kPlanItems];
    unsigned int rank[kPlanItems];
#pragma unroll
    for (int it = 0; it < kPlanItems; ++it) {
        const long long i = base + it * kPlanThreads + threadIdx.x;
        mm[it] = -1;
        if (i < n) {
            const int m = mid[i];
            mm[it] = (m >= 0 && m < M) ? m : M;
            rank[it] = atomicAdd(&h[mm[it]], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= M; i += kPlanThreads)
        start[i] = h[i] ? seg_off[i] + atomicAdd(&cursor[i], h[i]) : 0u;
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kPlanItems; ++it) {
        const long long i = base + it * kPlanThreads + threadIdx.x;
        if (mm[it] >= 0) perm[start[mm[it]] + rank[it]] = (unsigned int)i;
    }
}

// inactive rows (material id outside [0, M)): no kernel touches them, so their outputs are zeroed here
__global__ void multi_zero_inactive_kernel(long long n, const unsigned int* __restrict__ seg_off_inactive,
                                           const unsigned int* __restrict__ perm, float* __restrict__ out_dir, int dir_cols,
                                           float* __restrict__ out_pdf) {
    const unsigned int first = *seg_off_inactive;
    for (long long p = first + (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const unsigned int i = perm[p];
        if (out_dir) for (int c = 0; c < dir_cols; ++c) out_dir[(size_t)i * dir_cols + c] = 0.0f;
        if (out_pdf) out_pdf[i] = 0.0f;
    }
}

size_t multi_scratch_bytes(long long n, int M) {
    if (n < 0 || M < 1 || M > kMaxMaterials) return 0;
    const size_t tiles = (size_t)(n / 128 + M + 1);
    return sizeof(unsigned int) * (size_t)kPlanHeaderWords + sizeof(int4) * tiles + sizeof(unsigned int) * 2 * (size_t)n +
           sizeof(float) * 2 * (size_t)n;
}

MultiPlanView multi_view(void* scratch, long long n, int M) {
    MultiPlanView v;
    unsigned int* w = static_cast<unsigned int*>(scratch);
    v.counts = w; v.cursor = w + 256; v.seg_off = w + 512; v.n_tiles = w + 769; v.fix_count = w + 1024;
    v.tiles = reinterpret_cast<int4*>(w + kPlanHeaderWords);
    v.perm = reinterpret_cast<unsigned int*>(v.tiles + (size_t)(n / 128 + M + 1));
    v.fix_list = v.perm + n;
    v.x0 = reinterpret_cast<float*>(v.fix_list + n);
    return v;
}

int launch_multi_plan(long long n, const int* material_id, int M, void* scratch, cudaStream_t stream) {
    if (n < 0 || n > 0x7fffffffll || M < 1 || M > kMaxMaterials || !scratch || (n > 0 && !material_id)) return -1;
    MultiPlanView v = multi_view(scratch, n, M);
    if (cudaMemsetAsync(scratch, 0, sizeof(unsigned int) * kPlanHeaderWords, stream) != cudaSuccess) return -3;
    const long long per_block = kPlanThreads * kPlanItems;
    const unsigned int blocks = (unsigned int)((n + per_block - 1) / per_block);
    if (blocks) multi_count_kernel<<<blocks, kPlanThreads, 0, stream>>>(material_id, n, M, v.counts);
    const unsigned int tile_blocks = (unsigned int)((n / 128 + M + kPlanThreads - 1) / kPlanThreads);
    multi_tiles_kernel<<<tile_blocks < 1 ? 1 : (tile_blocks > 296 ? 296 : tile_blocks), kPlanThreads, 0, stream>>>(
        n, M, v.counts, v.seg_off, v.n_tiles, v.tiles);
    if (blocks) multi_scatter_kernel<<<blocks, kPlanThreads, 0, stream>>>(material_id, n, M, v.seg_off, v.cursor, v.perm);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

int launch_multi_zero_inactive(long long n, int M, void* scratch, float* out_dir, int dir_cols, float* out_pdf,
                               cudaStream_t stream) {
    MultiPlanView v = multi_view(scratch, n, M);
    multi_zero_inactive_kernel<<<64, 256, 0, stream>>>(n, v.seg_off + M, v.perm, out_dir, dir_cols, out_pdf);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace bsdfdiff
