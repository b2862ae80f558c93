// PREC_FP32 path for the 32-wide sampler nets, and the fix-up pass of the tensor-core path: the whole per-query flow in
// fp32 on CUDA cores with EIGHT LANES PER QUERY (lane s of a query's group owns hidden neurons 4s .. 4s+3).
//
// Why not one thread per query (flow_simt.cu, which remains for the 64-wide nets, the forward-only mode and T = 0):
// a query is ~30 000 dependent-ish instructions; one thread per query makes the latency of ANY launch that of a whole
// query on one thread (78 us on B200), however few rows there are -- and the fix-up pass of the tensor-core path
//  typically recomputes 0.04-0.5 % of a batch.  Eight lanes per query cut that chain to ~1/7 (the
// 32 x 96 FMA of a hidden layer become 32 x 12 per lane), four queries share a warp, 16 a CTA, and a CTA needs 22 KB of
// shared memory instead of 78 KB, so 8 CTAs fit an SM.  The same kernel runs the full-batch fp32 parity path, which makes
// a fixed-up row bit-identical to the fp32 result of that row by construction (tests/test_gpu_parity.py).
//
// Arithmetic: identical formulas to flow_simt.cu / the reference (rendering/utils/mlp_brdf_sampling.py:17-181); the sums
// over the hidden dimension are accumulated per lane over k = 0..31 in order (as flow_simt does), the 2-row output layer
// and the base net's output are reduced over the eight lanes with a butterfly.
#include "common.cuh"

namespace bsdfdiff {

#ifndef BSDFDIFF_L8_UNROLL
#define BSDFDIFF_L8_UNROLL 4
#endif
constexpr int kL8Unroll = BSDFDIFF_L8_UNROLL;   // k-blocks of the hidden-layer product per loop iteration (tuning builds)
constexpr int kL8Threads = 128;               // 4 warps x 4 queries
constexpr int kL8Rows = 16;                   // queries per CTA step
constexpr int kL8Act = 100;                   // floats per query row in shared memory: h[32] u[32] v[32] + 4 pad (bank shift 4 per row)
constexpr int kH = 32;

__device__ __forceinline__ float group_sum(float v, unsigned mask) {
    v += __shfl_xor_sync(mask, v, 1, 8);
    v += __shfl_xor_sync(mask, v, 2, 8);
    v += __shfl_xor_sync(mask, v, 4, 8);
    return v;
}
__device__ __forceinline__ void silu_grad8(float z, float& h, float& g) {
#ifdef BSDFDIFF_LANE8_IEEE
    const float s = sigmoid_precise(z);
#else
    const float s = sigmoid_newton(z);
#endif
    h = z * s;
    g = s * fmaf(z, 1.0f - s, 1.0f);          // silu'(z) = s (1 + z (1 - s))
}

// One query (row i of the caller's tensors) on the eight lanes of group `mask`; `sub` = this lane's index in the group,
// `act` = the group's row in shared memory.  All eight lanes carry the scalar state redundantly.
template <bool TANGENTS>
__device__ __forceinline__ void lane8_row(const FlowParams& P, const float* __restrict__ W, const float* __restrict__ base,
                                          float* __restrict__ act, int sub, unsigned mask, long long i,
                                          const float2* __restrict__ x0_fix) {
    const int k0 = (P.domain == kDisk) ? 3 : 4;      // first PE column of layer 1
    const float inv_t = (float)(1.0 / (double)P.T);
    const int j0 = 4 * sub;
    float w0, w1, wiz;
    load_wi(P, i, w0, w1, wiz);

    // ---- PE5(wi): lanes 0..4 evaluate one frequency each, everybody reads all 22 values back ---------------------------
    float e[kPE5];
    {
        if (sub < 5) {
            const float f = (float)(1 << sub);
            float s0, c0, s1, c1;
            sincosf(w0 * f, &s0, &c0);
            sincosf(w1 * f, &s1, &c1);
            *reinterpret_cast<float4*>(act + 4 + 4 * sub) = make_float4(s0, s1, c0, c1);     // e[2 + 4k ..] at act[4 + 4k ..]
        }
        __syncwarp(mask);
        e[0] = w0; e[1] = w1;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const float4 q = *reinterpret_cast<const float4*>(act + 4 + 4 * k);
            e[2 + 4 * k] = q.x; e[3 + 4 * k] = q.y; e[4 + 4 * k] = q.z; e[5 + 4 * k] = q.w;
        }
        __syncwarp(mask);
    }
    // layer-1 contribution of PE5(wi) to this lane's four neurons: constant over the T steps (model.py:494 recomputes it)
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (P.T > 0) {
#pragma unroll
        for (int k = 0; k < kPE5; ++k) {
            const float4 w = *reinterpret_cast<const float4*>(W + (k0 + k) * kH + j0);
            bias.x = fmaf(e[k], w.x, bias.x); bias.y = fmaf(e[k], w.y, bias.y);
            bias.z = fmaf(e[k], w.z, bias.z); bias.w = fmaf(e[k], w.w, bias.w);
        }
    }
    // ---- base net p = Wo silu(W1 PE3 + b1) + bo: two of the 16 neurons per lane, butterfly over the group ---------------
    float bp[4] = {0.f, 0.f, 0.f, 0.f};
    auto base_eval8 = [&]() {
        float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const int j = 2 * sub + jj;
            float z = base[224 + j];
#pragma unroll
            for (int k = 0; k < kPE3; ++k) z = fmaf(e[k], base[j * kPE3 + k], z);      // PE3 is a prefix of PE5
            const float h = z * sigmoid_precise(z);
            q[0] = fmaf(h, base[240 + j], q[0]); q[1] = fmaf(h, base[256 + j], q[1]);
            q[2] = fmaf(h, base[272 + j], q[2]); q[3] = fmaf(h, base[288 + j], q[3]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) bp[c] = group_sum(q[c], mask) + base[304 + c];
    };

    float x0, x1, R = 1.0f, p0 = 1.0f;
    float wox = 0.0f, woy = 0.0f, woz = 1.0f, theta_o = 0.0f;
    if (P.mode == kModePdf) {
        load_wo(P, i, x0, x1, wox, woy, woz);
        theta_o = x0;
    } else {
        if (base) base_eval8();
        if (x0_fix) {                                  // fix-up pass: the base sample the tensor-core kernel stored with the row
            const float2 t = *x0_fix;
            x0 = t.x; x1 = t.y;
        } else if (P.x0) {
            const float2 t = reinterpret_cast<const float2*>(P.x0)[i];
            x0 = t.x; x1 = t.y;
        } else {
            float un[3];
            if (P.u_noise) { un[0] = P.u_noise[3 * i]; un[1] = P.u_noise[3 * i + 1]; un[2] = P.u_noise[3 * i + 2]; }
            base_draw(P.domain, bp, P.seed, P.offset, P.first_index + i, x0, x1, P.u_noise ? un : nullptr);
        }
        if (P.out_x0 && sub == 0) reinterpret_cast<float2*>(P.out_x0)[i] = make_float2(x0, x1);
        if (P.mode == kModeSample) p0 = expf(base_logprob(P.domain, bp, x0, x1));
    }

    // first-layer weights of the state inputs for this lane's neurons
    const float4 wa = *reinterpret_cast<const float4*>(W + 0 * kH + j0), wb = *reinterpret_cast<const float4*>(W + 1 * kH + j0);
    const float4 wc = *reinterpret_cast<const float4*>(W + 2 * kH + j0), wd = *reinterpret_cast<const float4*>(W + 3 * kH + j0);
    const float* Wout = W + P.in_dim * kH + (P.n_hidden - 1) * kH * kH;     // [k][2]
    const float sgn = (P.mode == kModePdf) ? -1.0f : 1.0f;

    for (int t = 0; t < P.T; ++t) {
        // alpha = t/T (forward, mlp_brdf_sampling.py:27) or 1 - t/T (reverse, :78), in double then fp32
        const float alpha = (P.mode == kModePdf) ? (float)(1.0 - (double)t / (double)P.T) : (float)((double)t / (double)P.T);
        float s = 0.0f, c = 1.0f;
        if (P.domain == kSpherical) sincosf(x1, &s, &c);
        const float i0 = x0, i1 = (P.domain == kDisk) ? x1 : s, i2 = (P.domain == kDisk) ? alpha : c;
        // ---- layer 1: this lane's four neurons ------------------------------------------------------------------------
        {
            const float bz[4] = {bias.x, bias.y, bias.z, bias.w};
            const float a4[4] = {wa.x, wa.y, wa.z, wa.w}, b4[4] = {wb.x, wb.y, wb.z, wb.w};
            const float c4[4] = {wc.x, wc.y, wc.z, wc.w}, d4[4] = {wd.x, wd.y, wd.z, wd.w};
            float h[4], u[4], v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float z = bz[q];
                z = fmaf(i0, a4[q], z);
                z = fmaf(i1, b4[q], z);
                z = fmaf(i2, c4[q], z);
                if (P.domain == kSpherical) z = fmaf(alpha, d4[q], z);
                float g;
                silu_grad8(z, h[q], g);
                u[q] = g * a4[q];
                v[q] = g * ((P.domain == kDisk) ? b4[q] : fmaf(c, b4[q], -s * c4[q]));
            }
            *reinterpret_cast<float4*>(act + j0) = make_float4(h[0], h[1], h[2], h[3]);
            if (TANGENTS) {
                *reinterpret_cast<float4*>(act + kH + j0) = make_float4(u[0], u[1], u[2], u[3]);
                *reinterpret_cast<float4*>(act + 2 * kH + j0) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
        __syncwarp(mask);
        // ---- hidden layers: z = W h, (u, v) <- silu'(z) (W u, W v) --------------------------------------------------------
        const float* Wl = W + P.in_dim * kH;
        for (int l = 1; l < P.n_hidden; ++l) {
            float az[4] = {0.f, 0.f, 0.f, 0.f}, au[4] = {0.f, 0.f, 0.f, 0.f}, av[4] = {0.f, 0.f, 0.f, 0.f};
            // four k per iteration: the row's h / u / v arrive as one float4 each (7 shared-memory loads per 48 FMAs instead
            // of 16); the accumulation order over k is unchanged
#pragma unroll kL8Unroll
            for (int k4 = 0; k4 < kH; k4 += 4) {
                const float4 h4 = *reinterpret_cast<const float4*>(act + k4);
                float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f), v4 = u4;
                if (TANGENTS) {
                    u4 = *reinterpret_cast<const float4*>(act + kH + k4);
                    v4 = *reinterpret_cast<const float4*>(act + 2 * kH + k4);
                }
                const float hs[4] = {h4.x, h4.y, h4.z, h4.w}, us[4] = {u4.x, u4.y, u4.z, u4.w}, vs[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 w = *reinterpret_cast<const float4*>(Wl + (k4 + q) * kH + j0);
                    az[0] = fmaf(hs[q], w.x, az[0]); az[1] = fmaf(hs[q], w.y, az[1]); az[2] = fmaf(hs[q], w.z, az[2]); az[3] = fmaf(hs[q], w.w, az[3]);
                    if (TANGENTS) {
                        au[0] = fmaf(us[q], w.x, au[0]); au[1] = fmaf(us[q], w.y, au[1]); au[2] = fmaf(us[q], w.z, au[2]); au[3] = fmaf(us[q], w.w, au[3]);
                        av[0] = fmaf(vs[q], w.x, av[0]); av[1] = fmaf(vs[q], w.y, av[1]); av[2] = fmaf(vs[q], w.z, av[2]); av[3] = fmaf(vs[q], w.w, av[3]);
                    }
                }
            }
            __syncwarp(mask);                         // every lane has read the layer's inputs
            float h[4], g[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) silu_grad8(az[q], h[q], g[q]);
            *reinterpret_cast<float4*>(act + j0) = make_float4(h[0], h[1], h[2], h[3]);
            if (TANGENTS) {
                *reinterpret_cast<float4*>(act + kH + j0) = make_float4(g[0] * au[0], g[1] * au[1], g[2] * au[2], g[3] * au[3]);
                *reinterpret_cast<float4*>(act + 2 * kH + j0) = make_float4(g[0] * av[0], g[1] * av[1], g[2] * av[2], g[3] * av[3]);
            }
            __syncwarp(mask);
            Wl += kH * kH;
        }
        // ---- output layer (2 rows): four k's per lane, butterfly -----------------------------------------------------------
        float d[2] = {0.f, 0.f}, du[2] = {0.f, 0.f}, dv[2] = {0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = j0 + q;
            const float2 w = reinterpret_cast<const float2*>(Wout)[k];
            const float hk = act[k];
            d[0] = fmaf(hk, w.x, d[0]); d[1] = fmaf(hk, w.y, d[1]);
            if (TANGENTS) {
                const float uk = act[kH + k], vk = act[2 * kH + k];
                du[0] = fmaf(uk, w.x, du[0]); du[1] = fmaf(uk, w.y, du[1]);
                dv[0] = fmaf(vk, w.x, dv[0]); dv[1] = fmaf(vk, w.y, dv[1]);
            }
        }
        d[0] = group_sum(d[0], mask); d[1] = group_sum(d[1], mask);
        if (TANGENTS) {
            du[0] = group_sum(du[0], mask); du[1] = group_sum(du[1], mask);
            dv[0] = group_sum(dv[0], mask); dv[1] = group_sum(dv[1], mask);
            // J = I +- (1/T) dd/dx ; det = J00 J11 - J01 J10   (mlp_brdf_sampling.py:44-47 / 96-99)
            const float j00 = 1.0f + sgn * inv_t * du[0], j01 = sgn * inv_t * dv[0];
            const float j10 = sgn * inv_t * du[1], j11 = 1.0f + sgn * inv_t * dv[1];
            const float det = j00 * j11 - j01 * j10;
            R = (P.mode == kModePdf) ? R * det : R / det;
        }
        x0 = fmaf(sgn * inv_t, d[0], x0);
        x1 = fmaf(sgn * inv_t, d[1], x1);
        __syncwarp(mask);                             // the next step's first layer overwrites the row
    }

    if (P.mode == kModeSample) {
        if (sub == 0) store_sample(P, i, x0, x1, p0 * R);
    } else if (P.mode == kModePdf) {
        base_eval8();
        const float lp = base_logprob(P.domain, bp, x0, x1);
        if (sub == 0) {
            if (P.log_output) P.out_pdf[i] = lp;        // model.py:393-398 / 308-317 return the LOG density
            else store_pdf(P, i, expf(lp) * R, wiz, wox, woy, woz, theta_o);
        }
    }
}

template <bool TANGENTS>
#ifndef BSDFDIFF_L8_MINB
#define BSDFDIFF_L8_MINB 4
#endif
__global__ void __launch_bounds__(kL8Threads, BSDFDIFF_L8_MINB) flow_lane8_kernel(const FlowParams P) {
    extern __shared__ __align__(16) float smem[];
    const int n_w = f32_image_floats(P.in_dim, kH, P.n_hidden);
    float* W = smem;
    float* base = W + ((n_w + 3) & ~3);
    float* acts = base + ((kBaseFloats + 3) & ~3);
    const int tid = threadIdx.x, lane = tid & 31;
    const int sub = lane & 7, rl = lane >> 3;
    const int slot = (tid >> 5) * 4 + rl;                      // this group's query slot in the CTA step
    const unsigned mask = 0xFFu << (8 * rl);
    float* act = acts + slot * kL8Act;
    auto stage = [&](const unsigned char* flow, const float* bsrc) {
        const PackedHeader* hdr = reinterpret_cast<const PackedHeader*>(flow);
        const float* src = reinterpret_cast<const float*>(flow + hdr->off_f32);
        for (int i = tid; i < n_w; i += kL8Threads) W[i] = src[i];
        if (bsrc) for (int i = tid; i < kBaseFloats; i += kL8Threads) base[i] = bsrc[i];
    };

    // Iteration spaces (one call site of lane8_row for all of them: a row must not depend on the launch shape):
    //   single material:    rows / list entries  blockIdx.x * 16 + slot, grid-strided
    //   multi, main pass:   the plan's virtual tiles of <= 128 rows of one material, 16 rows per CTA step
    //   multi, fix-up pass: material m's flagged rows fix_list[seg_off[m] .. + fix_count[m]), 16 per CTA step
    const bool multi = P.n_materials > 0;
    int staged = -1;
    auto need = [&](int m) {
        if (m == staged) return;
        __syncthreads();
        stage(P.flows[m], P.bases[m]);
        __syncthreads();
        staged = m;
    };
    long long n_rows = P.n;
    if (!multi) {
        if (P.fix_pass) {
            n_rows = (long long)min(*P.fix_count, (unsigned int)min(P.n, (long long)0xffffffffll));
            if ((long long)blockIdx.x * kL8Rows >= n_rows) return;
        }
        stage(P.flow, P.base);
        __syncthreads();
    }
    const float* bptr = (multi || P.base) ? base : nullptr;
    const unsigned int n_work = (multi && !P.fix_pass) ? *P.n_tiles_dev * 8u : 0u;       // 8 CTA steps per 128-row tile
    long long jj = (long long)blockIdx.x * kL8Rows + slot;
    unsigned int k = blockIdx.x;
    int m = 0;
    const bool x0_listed = P.fix_pass && P.fix_x0 && P.mode == kModeSample;
    for (;;) {
        long long i = -1, at = 0;                        // at = position in the fix-up list (base samples are stored alongside)
        if (!multi) {
            if ((long long)(jj - slot) >= n_rows) break;         // uniform over the CTA
            if (jj < n_rows) { i = P.fix_pass ? (long long)P.fix_list[jj] : jj; at = jj; }
            jj += (long long)gridDim.x * kL8Rows;
        } else if (!P.fix_pass) {
            if (k >= n_work) break;
            const int4 ti = P.tiles[k >> 3];
            need(ti.x);
            const int r = (int)(k & 7u) * kL8Rows + slot;
            if (r < ti.z) i = (long long)P.perm[ti.y + r];
            k += gridDim.x;
        } else {
            while (m < P.n_materials && (unsigned long long)k * kL8Rows >= P.fix_count[m]) { ++m; k = blockIdx.x; }
            if (m >= P.n_materials) break;
            need(m);
            const unsigned int e = k * kL8Rows + slot;
            if (e < P.fix_count[m]) { at = (long long)P.seg_off[m] + e; i = (long long)P.fix_list[at]; }
            k += gridDim.x;
        }
        if (i >= 0) lane8_row<TANGENTS>(P, W, bptr, act, sub, mask, i,
                                        x0_listed ? reinterpret_cast<const float2*>(P.fix_x0) + at : nullptr);
    }
}

int launch_lane8(const FlowParams& P, cudaStream_t stream) {
    if (P.hidden != kH || P.T < 1 || P.mode == kModeForward || P.n_hidden < 2 || P.n_hidden > 8) return -2;
    if (!P.flow && P.n_materials <= 0) return -2;
    const int n_w = f32_image_floats(P.in_dim, kH, P.n_hidden);
    const size_t smem = sizeof(float) * (((n_w + 3) & ~3) + ((kBaseFloats + 3) & ~3) + (size_t)kL8Rows * kL8Act);
    auto kern = flow_lane8_kernel<true>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kL8Threads, smem) != cudaSuccess || occ < 1) return -3;
    if (P.fix_pass && (!P.fix_count || !P.fix_list)) return -1;
    long long work = (P.n_materials > 0) ? (P.n / 128 + P.n_materials) * 8 : (P.n + kL8Rows - 1) / kL8Rows;
    long long grid = (long long)sms * occ;
    if (grid > work) grid = work;
    if (grid < 1) return 0;
    kern<<<(unsigned)grid, kL8Threads, smem, stream>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace bsdfdiff
