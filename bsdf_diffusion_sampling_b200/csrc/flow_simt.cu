// PREC_FP32 path: the whole per-query flow (base sample / log-prob, T Euler steps of the velocity MLP
// with two forward-mode tangent columns, 2x2 determinant product, domain epilogue) on CUDA cores in
// fp32, one thread per query, weights resident in shared memory, state never leaves the SM.
//
// This is the parity path (matches the reference's fp32 eager results to ~1e-5; see
// tests/test_gpu_parity.py) and the fallback for shapes the tensor-core kernel does not cover.
// It replaces, per query, what the reference does with ~90 eager launches per Euler step
// (rendering/utils/mlp_brdf_sampling.py:26-47, 77-99, 116-136, 154-176).
//
// Data layout in shared memory (per CTA of kThreads queries):
//   weights  fp32 image of the packed blob ([k][j] input-major, see common.cuh)
//   base     308 floats
//   act      per-thread columns act[row*kThreads + tid]; rows [0,H) = h, [H,2H) = u = dh/dx0,
//            [2H,3H) = v = dh/dx1, [3H,4H) = first-layer contribution of PE5(wi) (constant over steps)
//            -> bank-conflict-free (consecutive threads hit consecutive banks), weight reads are
//            warp-wide broadcasts of float4.
#include "common.cuh"

namespace bsdfdiff {

constexpr int kThreads = 128;

template <int H, bool TANGENTS>
struct Act {
    float* col;  // this thread's column base (already offset by tid)
    __device__ __forceinline__ float& h(int k) { return col[k * kThreads]; }
    __device__ __forceinline__ float& u(int k) { return col[(H + k) * kThreads]; }
    __device__ __forceinline__ float& v(int k) { return col[(2 * H + k) * kThreads]; }
    __device__ __forceinline__ float& bias(int k) { return col[((TANGENTS ? 3 : 1) * H + k) * kThreads]; }
};

__device__ __forceinline__ void silu_grad(float z, float& h, float& g) {
    float s = sigmoid_precise(z);
    h = z * s;
    g = s * fmaf(z, 1.0f - s, 1.0f);          // silu'(z) = s (1 + z (1 - s))
}

// hidden layer: z = W h, (u,v) <- silu'(z) * (W u, W v); Wt is [k][j]
template <int H, bool TANGENTS>
__device__ __forceinline__ void hidden_layer(const float* __restrict__ Wt, Act<H, TANGENTS> a) {
    float az[H], au[TANGENTS ? H : 1], av[TANGENTS ? H : 1];
#pragma unroll
    for (int j = 0; j < H; ++j) { az[j] = 0.0f; if (TANGENTS) { au[j] = 0.0f; av[j] = 0.0f; } }
#pragma unroll 2
    for (int k = 0; k < H; ++k) {
        const float hk = a.h(k);
        float uk = 0.0f, vk = 0.0f;
        if (TANGENTS) { uk = a.u(k); vk = a.v(k); }
        const float4* w4 = reinterpret_cast<const float4*>(Wt + k * H);
#pragma unroll
        for (int j4 = 0; j4 < H / 4; ++j4) {
            const float4 w = w4[j4];
            az[4 * j4 + 0] = fmaf(hk, w.x, az[4 * j4 + 0]);
            az[4 * j4 + 1] = fmaf(hk, w.y, az[4 * j4 + 1]);
            az[4 * j4 + 2] = fmaf(hk, w.z, az[4 * j4 + 2]);
            az[4 * j4 + 3] = fmaf(hk, w.w, az[4 * j4 + 3]);
            if (TANGENTS) {
                au[4 * j4 + 0] = fmaf(uk, w.x, au[4 * j4 + 0]);
                au[4 * j4 + 1] = fmaf(uk, w.y, au[4 * j4 + 1]);
                au[4 * j4 + 2] = fmaf(uk, w.z, au[4 * j4 + 2]);
                au[4 * j4 + 3] = fmaf(uk, w.w, au[4 * j4 + 3]);
                av[4 * j4 + 0] = fmaf(vk, w.x, av[4 * j4 + 0]);
                av[4 * j4 + 1] = fmaf(vk, w.y, av[4 * j4 + 1]);
                av[4 * j4 + 2] = fmaf(vk, w.z, av[4 * j4 + 2]);
                av[4 * j4 + 3] = fmaf(vk, w.w, av[4 * j4 + 3]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float h, g;
        silu_grad(az[j], h, g);
        a.h(j) = h;
        if (TANGENTS) { a.u(j) = g * au[j]; a.v(j) = g * av[j]; }
    }
}

// One velocity evaluation d = D(x, alpha | wi) with tangents du = dd/dx0, dv = dd/dx1.
// Disk input [x0,x1,alpha,PE5] (model.py:494-495), spherical [theta,sin phi,cos phi,alpha,PE5]
// (mlp_brdf_sampling.py:119-121); the PE5 part of layer 1 is the precomputed a.bias().
template <int H, bool TANGENTS>
__device__ __forceinline__ void velocity(const float* __restrict__ W, int domain, int in_dim, int n_hidden,
                                         Act<H, TANGENTS> a, float x0, float x1, float alpha,
                                         float d[2], float du[2], float dv[2]) {
    float s = 0.0f, c = 1.0f;
    if (domain == kSpherical) sincosf(x1, &s, &c);
    const float i0 = x0, i1 = (domain == kDisk) ? x1 : s, i2 = (domain == kDisk) ? alpha : c;
    const float* w0 = W, *w1 = W + H, *w2 = W + 2 * H, *w3 = W + 3 * H;
#pragma unroll 8
    for (int j = 0; j < H; ++j) {
        float z = a.bias(j);
        z = fmaf(i0, w0[j], z);
        z = fmaf(i1, w1[j], z);
        z = fmaf(i2, w2[j], z);
        if (domain == kSpherical) z = fmaf(alpha, w3[j], z);
        float h, g;
        silu_grad(z, h, g);
        a.h(j) = h;
        if (TANGENTS) {
            a.u(j) = g * w0[j];
            a.v(j) = g * ((domain == kDisk) ? w1[j] : fmaf(c, w1[j], -s * w2[j]));
        }
    }
    const float* Wl = W + in_dim * H;
    for (int l = 1; l < n_hidden; ++l) {
        hidden_layer<H, TANGENTS>(Wl, a);
        Wl += H * H;
    }
    d[0] = d[1] = du[0] = du[1] = dv[0] = dv[1] = 0.0f;
#pragma unroll 8
    for (int k = 0; k < H; ++k) {
        const float2 w = reinterpret_cast<const float2*>(Wl)[k];
        const float hk = a.h(k);
        d[0] = fmaf(hk, w.x, d[0]); d[1] = fmaf(hk, w.y, d[1]);
        if (TANGENTS) {
            const float uk = a.u(k), vk = a.v(k);
            du[0] = fmaf(uk, w.x, du[0]); du[1] = fmaf(uk, w.y, du[1]);
            dv[0] = fmaf(vk, w.x, dv[0]); dv[1] = fmaf(vk, w.y, dv[1]);
        }
    }
}

// The whole flow of one query (row i of the caller's tensors) with the staged weight set.
template <int H, bool TANGENTS>
__device__ __forceinline__ void simt_row(const FlowParams& P, const float* __restrict__ W, const float* __restrict__ base,
                                         Act<H, TANGENTS> a, long long i, const float2* __restrict__ x0_fix) {
    const int k0 = (P.domain == kDisk) ? 3 : 4;      // first PE column of layer 1
    const float inv_t = (float)(1.0 / (double)P.T);
    float w0, w1, wiz;
    load_wi(P, i, w0, w1, wiz);

    // layer-1 contribution of PE5(wi): constant over the T steps (the reference recomputes it
    // every step, model.py:494)
    if (P.T > 0) {
        float e[kPE5];
        positional_encoding<5>(w0, w1, e);
#pragma unroll 4
        for (int j = 0; j < H; ++j) {
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < kPE5; ++k) acc = fmaf(e[k], W[(k0 + k) * H + j], acc);
            a.bias(j) = acc;
        }
    }

    float x0, x1, R = 1.0f, p0 = 1.0f;
    float wox = 0.0f, woy = 0.0f, woz = 1.0f, theta_o = 0.0f;
    float bp[4] = {0.f, 0.f, 0.f, 0.f};
    if (P.mode == kModePdf) {
        load_wo(P, i, x0, x1, wox, woy, woz);
        theta_o = x0;
    } else {
        if (base) base_eval(base, w0, w1, bp);
        if (x0_fix) {                                  // fix-up pass: the base sample stored with the flagged row
            const float2 t = *x0_fix;
            x0 = t.x; x1 = t.y;
        } else if (P.x0) {
            float2 t = reinterpret_cast<const float2*>(P.x0)[i];
            x0 = t.x; x1 = t.y;
        } else {
            float un[3];
            if (P.u_noise) { un[0] = P.u_noise[3 * i]; un[1] = P.u_noise[3 * i + 1]; un[2] = P.u_noise[3 * i + 2]; }
            base_draw(P.domain, bp, P.seed, P.offset, P.first_index + i, x0, x1, P.u_noise ? un : nullptr);
        }
        if (P.out_x0) reinterpret_cast<float2*>(P.out_x0)[i] = make_float2(x0, x1);
        if (P.mode == kModeSample) p0 = expf(base_logprob(P.domain, bp, x0, x1));
    }

    const float sgn = (P.mode == kModePdf) ? -1.0f : 1.0f;
    for (int t = 0; t < P.T; ++t) {
        // alpha = t/T (forward, mlp_brdf_sampling.py:27) or 1 - t/T (reverse, :78), in double then fp32
        const float alpha = (P.mode == kModePdf) ? (float)(1.0 - (double)t / (double)P.T)
                                                 : (float)((double)t / (double)P.T);
        float d[2], du[2], dv[2];
        velocity<H, TANGENTS>(W, P.domain, P.in_dim, P.n_hidden, a, x0, x1, alpha, d, du, dv);
        if (TANGENTS) {
            // J = I +- (1/T) dd/dx ; det = J00 J11 - J01 J10   (mlp_brdf_sampling.py:44-47 / 96-99)
            const float j00 = 1.0f + sgn * inv_t * du[0], j01 = sgn * inv_t * dv[0];
            const float j10 = sgn * inv_t * du[1], j11 = 1.0f + sgn * inv_t * dv[1];
            const float det = j00 * j11 - j01 * j10;
            R = (P.mode == kModePdf) ? R * det : R / det;
        }
        x0 = fmaf(sgn * inv_t, d[0], x0);
        x1 = fmaf(sgn * inv_t, d[1], x1);
    }

    if (P.mode == kModeSample) {
        store_sample(P, i, x0, x1, p0 * R);
    } else if (P.mode == kModePdf) {
        base_eval(base, w0, w1, bp);
        const float lp = base_logprob(P.domain, bp, x0, x1);
        if (P.log_output) P.out_pdf[i] = lp;            // model.py:393-398 / 308-317 return the LOG density
        else store_pdf(P, i, expf(lp) * R, wiz, wox, woy, woz, theta_o);
    } else {
        reinterpret_cast<float2*>(P.out_dir)[i] = make_float2(x0, x1);
    }
}

template <int H, bool TANGENTS>
__global__ void __launch_bounds__(kThreads) flow_simt_kernel(const FlowParams P) {
    extern __shared__ __align__(16) float smem[];
    const int n_w = (P.flow || P.n_materials > 0) ? f32_image_floats(P.in_dim, H, P.n_hidden) : 0;     // T == 0: base net only
    float* W = smem;
    float* base = W + ((n_w + 3) & ~3);
    float* act = base + ((kBaseFloats + 3) & ~3);
    Act<H, TANGENTS> a{act + threadIdx.x};
    auto stage = [&](const unsigned char* flow, const float* bsrc) {
        if (flow) {
            const PackedHeader* hdr = reinterpret_cast<const PackedHeader*>(flow);
            const float* src = reinterpret_cast<const float*>(flow + hdr->off_f32);
            for (int i = threadIdx.x; i < n_w; i += kThreads) W[i] = src[i];
        }
        if (bsrc) for (int i = threadIdx.x; i < kBaseFloats; i += kThreads) base[i] = bsrc[i];
    };

    // Three iteration spaces share ONE call site of simt_row (a second inlined copy could be contracted differently by
    // the compiler, and a row must not depend on which launch shape computed it):
    //   single material:        rows blockIdx.x * 128 + tid, grid-strided (fix-up pass: entries of fix_list)
    //   multi, main pass:       virtual tiles of <= 128 rows of ONE material (multi.cu built the plan); the weight set in
    //                           shared memory is swapped when the tile's material differs
    //   multi, fix-up pass:     material m's flagged rows are fix_list[seg_off[m] .. + fix_count[m]), 128 per chunk
    // In the multi modes all threads of the CTA walk the same tile / chunk sequence, so the barriers in need() are uniform.
    const bool multi = P.n_materials > 0;
    int staged = -1;
    auto need = [&](int m) {
        if (m == staged) return;
        __syncthreads();                                   // everyone is done with the previous weight set
        stage(P.flows[m], P.bases[m]);
        __syncthreads();
        staged = m;
    };
    long long n_rows = P.n;
    if (!multi) {
        // fix-up pass: the rows to recompute are fix_list[0 .. *fix_count) (written by the tensor-core kernel that ran
        // before this launch on the same stream); most launches find an empty or tiny list and leave before staging weights
        if (P.fix_pass) {
            n_rows = (long long)min(*P.fix_count, (unsigned int)min(P.n, (long long)0xffffffffll));
            if ((long long)blockIdx.x * kThreads >= n_rows) return;
        }
        stage(P.flow, P.base);
        __syncthreads();
    }
    const float* bptr = (multi || P.base) ? base : nullptr;
    const unsigned int n_tiles = (multi && !P.fix_pass) ? *P.n_tiles_dev : 0u;
    long long jj = (long long)blockIdx.x * kThreads + threadIdx.x;       // single
    unsigned int k = blockIdx.x;                                         // multi main: tile; multi fix: chunk of material m
    int m = 0;
    const bool x0_listed = P.fix_pass && P.fix_x0 && P.mode == kModeSample;
    for (;;) {
        long long i = -1, at = 0;
        if (!multi) {
            if (jj >= n_rows) break;
            i = P.fix_pass ? (long long)P.fix_list[jj] : jj;
            at = jj;
            jj += (long long)gridDim.x * kThreads;
        } else if (!P.fix_pass) {
            if (k >= n_tiles) break;
            const int4 ti = P.tiles[k];
            need(ti.x);
            if ((int)threadIdx.x < ti.z) i = (long long)P.perm[ti.y + threadIdx.x];
            k += gridDim.x;
        } else {
            while (m < P.n_materials && (unsigned long long)k * kThreads >= P.fix_count[m]) { ++m; k = blockIdx.x; }
            if (m >= P.n_materials) break;
            need(m);
            const unsigned int e = k * kThreads + threadIdx.x;
            if (e < P.fix_count[m]) { at = (long long)P.seg_off[m] + e; i = (long long)P.fix_list[at]; }
            k += gridDim.x;
        }
        if (i >= 0) simt_row<H, TANGENTS>(P, W, bptr, a, i,
                                          x0_listed ? reinterpret_cast<const float2*>(P.fix_x0) + at : nullptr);
    }
}

template <int H, bool TANGENTS>
static int launch_simt_t(const FlowParams& P, cudaStream_t stream) {
    const int n_w = (P.flow || P.n_materials > 0) ? f32_image_floats(P.in_dim, H, P.n_hidden) : 0;
    const int rows = (TANGENTS ? 4 : 2) * H;
    const size_t smem = sizeof(float) * (((n_w + 3) & ~3) + ((kBaseFloats + 3) & ~3) + (size_t)rows * kThreads);
    auto kern = flow_simt_kernel<H, TANGENTS>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem) != cudaSuccess || occ < 1) return -3;
    long long tiles = (P.n_materials > 0) ? P.n / kThreads + P.n_materials : (P.n + kThreads - 1) / kThreads;
    long long grid = (long long)sms * occ;
    if (grid > tiles) grid = tiles;
    if (grid < 1) return 0;
    if (P.fix_pass && (!P.fix_count || !P.fix_list)) return -1;
    kern<<<(unsigned)grid, kThreads, smem, stream>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

int launch_simt(const FlowParams& P, cudaStream_t stream) {
    const bool tang = (P.mode != kModeForward);
    if (P.hidden == 32) return tang ? launch_simt_t<32, true>(P, stream) : launch_simt_t<32, false>(P, stream);
    if (P.hidden == 64) return tang ? launch_simt_t<64, true>(P, stream) : launch_simt_t<64, false>(P, stream);
    return -2;
}

// ---------------------------------------------------------------------------------------------
// Plain MLP forward (tinycudann.Network.forward replacement): out[n,2] = MLP(in[n,in_dim]), fp32.
// ---------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(kThreads) mlp_forward_kernel(long long n, const float* __restrict__ in, int in_dim,
                                                               const unsigned char* __restrict__ flow, int n_hidden,
                                                               float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    const PackedHeader* hdr = reinterpret_cast<const PackedHeader*>(flow);
    const int n_w = f32_image_floats(in_dim, H, n_hidden);
    float* W = smem;
    float* act = W + ((n_w + 3) & ~3);
    const float* src = reinterpret_cast<const float*>(flow + hdr->off_f32);
    for (int i = threadIdx.x; i < n_w; i += kThreads) W[i] = src[i];
    __syncthreads();
    Act<H, false> a{act + threadIdx.x};
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
        float z[H];
#pragma unroll
        for (int j = 0; j < H; ++j) z[j] = 0.0f;
        for (int k = 0; k < in_dim; ++k) {
            const float xk = in[i * in_dim + k];
#pragma unroll
            for (int j = 0; j < H; ++j) z[j] = fmaf(xk, W[k * H + j], z[j]);
        }
#pragma unroll
        for (int j = 0; j < H; ++j) a.h(j) = z[j] * sigmoid_precise(z[j]);
        const float* Wl = W + in_dim * H;
        for (int l = 1; l < n_hidden; ++l) { hidden_layer<H, false>(Wl, a); Wl += H * H; }
        float d0 = 0.0f, d1 = 0.0f;
        for (int k = 0; k < H; ++k) {
            const float2 w = reinterpret_cast<const float2*>(Wl)[k];
            d0 = fmaf(a.h(k), w.x, d0); d1 = fmaf(a.h(k), w.y, d1);
        }
        reinterpret_cast<float2*>(out)[i] = make_float2(d0, d1);
    }
}

template <int H>
static int launch_mlp_t(long long n, const float* in, int in_dim, const unsigned char* flow, int n_hidden, float* out,
                        cudaStream_t stream) {
    const int n_w = f32_image_floats(in_dim, H, n_hidden);
    const size_t smem = sizeof(float) * (((n_w + 3) & ~3) + (size_t)2 * H * kThreads);
    auto kern = mlp_forward_kernel<H>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem) != cudaSuccess || occ < 1) return -3;
    long long tiles = (n + kThreads - 1) / kThreads, grid = (long long)sms * occ;
    if (grid > tiles) grid = tiles;
    if (grid < 1) return 0;
    kern<<<(unsigned)grid, kThreads, smem, stream>>>(n, in, in_dim, flow, n_hidden, out);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

int launch_mlp_forward_simt(long long n, const float* in, int in_dim, const unsigned char* flow, int H, int n_hidden,
                            float* out, cudaStream_t stream) {
    if (H == 32) return launch_mlp_t<32>(n, in, in_dim, flow, n_hidden, out, stream);
    if (H == 64) return launch_mlp_t<64>(n, in, in_dim, flow, n_hidden, out, stream);
    return -2;
}

}  // namespace bsdfdiff
