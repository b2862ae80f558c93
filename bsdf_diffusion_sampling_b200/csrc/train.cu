// Training step of the flow nets: fused forward + backward + Adam for the flow-matching regression of the
// reference's diffusion / rectify stages -- one launch per optimisation step.
//
//   reference (learning_repo_cleanup/disk_domain_sampling.py:49-58, 123-131; spherical_domain_sampling.py:59-75):
//       alpha   = linspace(0, 1, N)
//       x_alpha = (1 - alpha) x_0 + alpha omega_o          (spherical: omega_o.phi first moved to within pi of x_0.phi,
//                                                            net input (theta, sin phi, cos phi))
//       pred    = D(x_alpha, alpha, omega_i)                (NN_cond_pos_simpler / NN_cond_pos / ..._complicate: bias-free SiLU MLP
//                                                            on [x.., alpha, PE5(omega_i)])
//       loss    = mean((pred - (omega_o - x_0))^2);  loss.backward();  Adam.step()
//   which is ~40 eager launches, an autograd graph with every [N, H] activation in HBM (N = 4.9 M rows: ~0.6 GB per
//   tensor) and a separate optimizer pass.  tiny-cuda-nn's analogue is kernel_mlp_fused_backward
//   (tiny-cuda-nn/src/fully_fused_mlp.cu:150-259) + its own optimizer kernels; the reference's scripts do not use it for
//   training.
//
// Here: persistent CTAs, a tile of ROWS rows per CTA iteration, fp32 throughout (the optimiser's master weights).
//   forward   thread <-> row; weights transposed in shared memory (warp-broadcast float4 reads); pre-activations z_l of
//             every layer stay in per-thread shared-memory columns (nothing goes to HBM)
//   backward  thread <-> row for the activation gradients (delta_h = W^T delta_z as broadcast reads of the same image);
//             the weight gradients dW_l = delta_z_l^T h_{l-1} are a [H x ROWS] . [ROWS x K] product over the tile, computed
//             cooperatively from the shared-memory columns and added to the global gradient with one atomic per weight
//             per tile
//   Adam      the last CTA to finish (device-side ticket) applies torch.optim.Adam's update to all parameters, clears
//             the gradient buffer and the ticket -- no second launch, no host synchronisation, CUDA-graph capturable.
#include "common.cuh"
#include "train.cuh"

#include <cstdlib>

namespace bsdfdiff {

__device__ __forceinline__ void silu_both(float z, float& h, float& g) {
    const float s = sigmoid_newton(z);
    h = z * s;
    g = s * fmaf(z, 1.0f - s, 1.0f);
}

template <int H, int ROWS>
struct TrainSmem {
    // per-thread columns, element e of row r at col[e * RS + r].  RS = ROWS + 4: with RS = ROWS every column would start
    // in bank 0 and the float4 reads of the weight-gradient product (different columns, same rows) would be 8-way
    // bank conflicts; + 4 floats shifts consecutive columns by one float4
    static constexpr int RS = ROWS + 4;
    static constexpr int kIn = 28;                       // first-layer input (25 | 26), padded
    __host__ __device__ static size_t floats(int in_dim, int n_hidden) {
        const size_t w = (size_t)in_dim * H + (size_t)(n_hidden - 1) * H * H + 2 * H;
        return ((w + 3) & ~(size_t)3) + (size_t)RS * ((size_t)n_hidden * H + kIn + 2 * H);
    }
};

// acc[b] (+)= sum_r D[j][r] * P[k0 + b][r] over the tile, for the (j, block of 4 consecutive k) pairs dealt to this thread:
// pair o = threadIdx.x + s * ROWS, s < SLOTS.  D: [J][RS], P: [K][RS] (shared-memory columns).  The sums stay in
// registers across all tiles of the CTA and reach the global gradient once, at the end (flush_outer).
// Work item o of a layer's weight-gradient product: output row j = o / kb and the four columns k = kq + b * kb
// (kq = o % kb, kb = ceil(K / 4), b = 0..3).  Strided columns: the threads of a warp then read CONSECUTIVE columns of P,
// which the RS = ROWS + 4 padding spreads over all banks (blocks of 4 adjacent columns would be 4-way conflicts).
template <int ROWS, int RS>
__device__ __forceinline__ void tile_outer_one(const float* __restrict__ D, const float* __restrict__ P, int K, int o,
                                               float* acc) {
    const int kb = (K + 3) >> 2;
    const int j = o / kb, kq = o - j * kb;
    float a0 = acc[0], a1 = acc[1], a2 = acc[2], a3 = acc[3];
    const float4* d4 = reinterpret_cast<const float4*>(D + (size_t)j * RS);
    const float4* p0 = reinterpret_cast<const float4*>(P + (size_t)kq * RS);
    const float4* p1 = reinterpret_cast<const float4*>(P + (size_t)min(kq + kb, K - 1) * RS);
    const float4* p2 = reinterpret_cast<const float4*>(P + (size_t)min(kq + 2 * kb, K - 1) * RS);
    const float4* p3 = reinterpret_cast<const float4*>(P + (size_t)min(kq + 3 * kb, K - 1) * RS);
#pragma unroll 4
    for (int r = 0; r < ROWS / 4; ++r) {
        const float4 d = d4[r];
        float4 p = p0[r];
        a0 = fmaf(d.x, p.x, a0); a0 = fmaf(d.y, p.y, a0); a0 = fmaf(d.z, p.z, a0); a0 = fmaf(d.w, p.w, a0);
        p = p1[r];
        a1 = fmaf(d.x, p.x, a1); a1 = fmaf(d.y, p.y, a1); a1 = fmaf(d.z, p.z, a1); a1 = fmaf(d.w, p.w, a1);
        p = p2[r];
        a2 = fmaf(d.x, p.x, a2); a2 = fmaf(d.y, p.y, a2); a2 = fmaf(d.z, p.z, a2); a2 = fmaf(d.w, p.w, a2);
        p = p3[r];
        a3 = fmaf(d.x, p.x, a3); a3 = fmaf(d.y, p.y, a3); a3 = fmaf(d.z, p.z, a3); a3 = fmaf(d.w, p.w, a3);
    }
    acc[0] = a0; acc[1] = a1; acc[2] = a2; acc[3] = a3;
}
// acc[0..3] of work item o -> the global gradient (row-major [J][K], torch layout)
__device__ __forceinline__ void add_outer_one(int K, int o, const float* acc, float* __restrict__ grad) {
    const int kb = (K + 3) >> 2;
    const int j = o / kb, kq = o - j * kb;
#pragma unroll
    for (int b = 0; b < 4; ++b)
        if (kq + b * kb < K) atomicAdd(grad + (size_t)j * K + kq + b * kb, acc[b]);
}
template <int ROWS, int RS, int SLOTS>
__device__ __forceinline__ void tile_outer(const float* __restrict__ D, int J, const float* __restrict__ P, int K,
                                           float (*acc)[4]) {
    const int kb = (K + 3) >> 2;
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) {
        const int o = threadIdx.x + sl * ROWS;
        if (o >= J * kb) break;
        tile_outer_one<ROWS, RS>(D, P, K, o, acc[sl]);
    }
}
template <int ROWS, int SLOTS>
__device__ __forceinline__ void flush_outer(int J, int K, const float (*acc)[4], float* __restrict__ grad) {
    const int kb = (K + 3) >> 2;
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) {
        const int o = threadIdx.x + sl * ROWS;
        if (o >= J * kb) break;
        add_outer_one(K, o, acc[sl], grad);
    }
}

// Loss accumulation + torch.optim.Adam by the last CTA to finish (device-side ticket): shared by both training kernels.
template <int THREADS>
__device__ __forceinline__ void finish_step(const TrainParams& P, int n_w, float loss_part) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss_part += __shfl_xor_sync(0xffffffffu, loss_part, o);
    if ((tid & 31) == 0 && loss_part != 0.0f) atomicAdd(P.loss, loss_part);
    __threadfence();
    __syncthreads();
    __shared__ unsigned int s_last;
    if (tid == 0) s_last = (atomicAdd(P.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (P.apply_update) {
        // torch.optim.Adam (amsgrad = False, weight_decay = 0, maximize = False)
        const float bc1 = 1.0f - powf(P.beta1, (float)P.step), bc2 = 1.0f - powf(P.beta2, (float)P.step);
        const float step_size = P.lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
        for (int i = tid; i < n_w; i += THREADS) {
            const float g = __ldcg(P.grad + i);
            const float m = P.beta1 * P.adam_m[i] + (1.0f - P.beta1) * g;
            const float v = P.beta2 * P.adam_v[i] + (1.0f - P.beta2) * g * g;
            P.adam_m[i] = m; P.adam_v[i] = v;
            P.weights[i] -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + P.eps);
            P.grad[i] = 0.0f;
        }
    }
    if (tid == 0) *P.ticket = 0u;
}

template <int H, int ROWS>
__global__ void __launch_bounds__(ROWS) flow_matching_step_kernel(const TrainParams P) {
    extern __shared__ __align__(16) float smem[];
    const int in_dim = P.in_dim, L = P.n_hidden;
    const int n_w = in_dim * H + (L - 1) * H * H + 2 * H;
    constexpr int RS = TrainSmem<H, ROWS>::RS;
    constexpr int SLOTS = (H * (H / 4) + ROWS - 1) / ROWS;       // (j, 4-k block) pairs of one layer's gradient per thread
    // H = 32: the gradient sums of up to 4 hidden layers stay in registers for the CTA's lifetime (28 accumulators per
    // thread); H = 64 would need 384, so there every tile adds its product to the global gradient directly
    constexpr bool kRegAcc = (H == 32);
    constexpr int kMaxL = kRegAcc ? 4 : 1;
    float* Wt = smem;                                    // transposed image: W1t [in][H], W2t.. [H][H], Woutt [H][2]
    float* col = smem + ((n_w + 3) & ~3);
    float* Z = col;                                      // [L][H][RS]
    float* IN = Z + (size_t)L * H * RS;                  // [kIn][RS]
    float* D = IN + (size_t)TrainSmem<H, ROWS>::kIn * RS;   // [H][RS]  delta_z of the current layer
    float* HP = D + (size_t)H * RS;                      // [H][RS]  input of the current layer (h_{l-1})
    const int tid = threadIdx.x;
    float gacc[kMaxL][SLOTS][4], gout[1][4];             // this thread's share of dW_1..dW_L and dWout, summed over the CTA's tiles
#pragma unroll
    for (int l = 0; l < kMaxL; ++l)
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl) gacc[l][sl][0] = gacc[l][sl][1] = gacc[l][sl][2] = gacc[l][sl][3] = 0.0f;
    gout[0][0] = gout[0][1] = gout[0][2] = gout[0][3] = 0.0f;

    {   // stage the weights, transposed (torch layout [out][in] -> [in][out])
        const float* w = P.weights;
        float* t = Wt;
        for (int l = 0; l <= L; ++l) {
            const int R = (l == L) ? 2 : H, C = (l == 0) ? in_dim : H;
            for (int i = tid; i < R * C; i += ROWS) { const int j = i / C, k = i - j * C; t[k * R + j] = w[i]; }
            w += R * C; t += R * C;
        }
    }
    __syncthreads();

    const long long n = P.n;
    const long long n_tiles = (n + ROWS - 1) / ROWS;
    const float inv_nm1 = (n > 1) ? 1.0f / (float)(n - 1) : 0.0f;
    const float inv_n = 1.0f / (float)n;                 // d mean((pred - target)^2 over [N,2]) / d pred = (pred - target) / N
    float loss_acc = 0.0f;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long i = tile * ROWS + tid;
        const bool valid = i < n;
        const long long ic = valid ? i : n - 1;
        // ---- inputs ----------------------------------------------------------------------------------------
        const float2 a = reinterpret_cast<const float2*>(P.x0)[ic];
        float2 b = reinterpret_cast<const float2*>(P.x1)[ic];
        const float2 wi = reinterpret_cast<const float2*>(P.wi)[ic];
        const float alpha = P.alpha ? P.alpha[ic] : (float)ic * inv_nm1;       // torch.linspace(0, 1, N)
        float t0 = b.x - a.x, t1 = b.y - a.y;            // regression target omega_o - x_0
        int k0;
        if (P.domain == kDisk) {
            IN[0 * RS + tid] = (1.0f - alpha) * a.x + alpha * b.x;
            IN[1 * RS + tid] = (1.0f - alpha) * a.y + alpha * b.y;
            IN[2 * RS + tid] = alpha;
            k0 = 3;
        } else {
            // spherical_domain_sampling.py:61-72: move omega_o.phi to within pi of x_0.phi, interpolate, embed periodically
            const float kPi = 3.14159265358979323846f, kTwoPi = 6.28318530717958647692f;
            if (t1 < -kPi) { b.y += kTwoPi; t1 += kTwoPi; } else if (t1 > kPi) { b.y -= kTwoPi; t1 -= kTwoPi; }
            const float th = (1.0f - alpha) * a.x + alpha * b.x, ph = (1.0f - alpha) * a.y + alpha * b.y;
            float s, c;
            sincosf(ph, &s, &c);
            IN[0 * RS + tid] = th; IN[1 * RS + tid] = s; IN[2 * RS + tid] = c; IN[3 * RS + tid] = alpha;
            k0 = 4;
        }
        {
            float e[kPE5];
            positional_encoding<5>(wi.x, wi.y, e);
#pragma unroll
            for (int k = 0; k < kPE5; ++k) IN[(k0 + k) * RS + tid] = e[k];
        }

        // ---- forward ---------------------------------------------------------------------------------------
        const float* Wl = Wt;
        for (int l = 0; l < L; ++l) {
            const int K = (l == 0) ? in_dim : H;
            float az[H];
#pragma unroll
            for (int j = 0; j < H; ++j) az[j] = 0.0f;
            const float* zin = Z + (size_t)(l - 1) * H * RS;
            for (int k = 0; k < K; ++k) {
                float hk;
                if (l == 0) hk = IN[k * RS + tid];
                else { const float z = zin[k * RS + tid]; hk = z * sigmoid_newton(z); }
                const float4* w4 = reinterpret_cast<const float4*>(Wl + k * H);
#pragma unroll
                for (int j4 = 0; j4 < H / 4; ++j4) {
                    const float4 w = w4[j4];
                    az[4 * j4 + 0] = fmaf(hk, w.x, az[4 * j4 + 0]);
                    az[4 * j4 + 1] = fmaf(hk, w.y, az[4 * j4 + 1]);
                    az[4 * j4 + 2] = fmaf(hk, w.z, az[4 * j4 + 2]);
                    az[4 * j4 + 3] = fmaf(hk, w.w, az[4 * j4 + 3]);
                }
            }
            float* zo = Z + (size_t)l * H * RS;
#pragma unroll
            for (int j = 0; j < H; ++j) zo[j * RS + tid] = az[j];
            Wl += K * H;
        }
        // output layer + loss; HP <- h_L, D rows 0..1 <- delta_out
        float p0 = 0.0f, p1 = 0.0f;
        {
            const float* zl = Z + (size_t)(L - 1) * H * RS;
            for (int k = 0; k < H; ++k) {
                const float z = zl[k * RS + tid];
                const float h = z * sigmoid_newton(z);
                HP[k * RS + tid] = h;
                const float2 w = reinterpret_cast<const float2*>(Wl)[k];
                p0 = fmaf(h, w.x, p0); p1 = fmaf(h, w.y, p1);
            }
        }
        float d0 = valid ? (p0 - t0) : 0.0f, d1 = valid ? (p1 - t1) : 0.0f;
        loss_acc += 0.5f * (d0 * d0 + d1 * d1);
        d0 *= inv_n; d1 *= inv_n;
        D[0 * RS + tid] = d0; D[1 * RS + tid] = d1;
        __syncthreads();

        // ---- backward ----------------------------------------------------------------------------------------
        tile_outer<ROWS, RS, 1>(D, 2, HP, H, gout);                       // dWout [2][H]
        // delta_z of the last hidden layer
        float dz[H];
        {
            const float* zl = Z + (size_t)(L - 1) * H * RS;
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const float2 w = reinterpret_cast<const float2*>(Wl)[k];
                float h, g;
                silu_both(zl[k * RS + tid], h, g);
                dz[k] = (d0 * w.x + d1 * w.y) * g;
            }
        }
        for (int l = L - 1; l >= 0; --l) {
            const int K = (l == 0) ? in_dim : H;
            const int off = (l == 0) ? 0 : in_dim * H + (l - 1) * H * H;
            __syncthreads();                              // everyone is done reading D / HP of the layer above
#pragma unroll
            for (int j = 0; j < H; ++j) D[j * RS + tid] = dz[j];
            const float* src = (l == 0) ? IN : Z + (size_t)(l - 1) * H * RS;
            if (l > 0) {
                for (int k = 0; k < H; ++k) { const float z = src[k * RS + tid]; HP[k * RS + tid] = z * sigmoid_newton(z); }
            }
            __syncthreads();
            // dW_l [H][K]
            if (kRegAcc) {                                // the layer index selects the register block at compile time
#pragma unroll
                for (int q = 0; q < kMaxL; ++q)
                    if (q == l) tile_outer<ROWS, RS, SLOTS>(D, H, (l == 0) ? IN : HP, K, gacc[q]);
            } else {
#pragma unroll 1
                for (int sl0 = 0; sl0 < SLOTS; ++sl0) {   // one slot at a time through a 4-float scratch block
                    float t4[1][4] = {{0.0f, 0.0f, 0.0f, 0.0f}};
                    const int o = tid + sl0 * ROWS, kb = (K + 3) >> 2;
                    if (o >= H * kb) break;
                    tile_outer_one<ROWS, RS>(D, (l == 0) ? IN : HP, K, o, t4[0]);
                    add_outer_one(K, o, t4[0], P.grad + off);
                }
            }
            if (l > 0) {
                // delta_h_{l-1}[k] = sum_j W_l[j][k] delta_z_l[j] = sum_j Wt_l[k][j] dz[j];  delta_z_{l-1} = delta_h * silu'(z_{l-1})
                const float* Wc = Wt + off;
                float nz[H];
#pragma unroll 4
                for (int k = 0; k < H; ++k) {
                    const float4* w4 = reinterpret_cast<const float4*>(Wc + k * H);
                    float acc = 0.0f;
#pragma unroll
                    for (int j4 = 0; j4 < H / 4; ++j4) {
                        const float4 w = w4[j4];
                        acc = fmaf(w.x, dz[4 * j4 + 0], acc); acc = fmaf(w.y, dz[4 * j4 + 1], acc);
                        acc = fmaf(w.z, dz[4 * j4 + 2], acc); acc = fmaf(w.w, dz[4 * j4 + 3], acc);
                    }
                    float h, g;
                    silu_both(src[k * RS + tid], h, g);
                    nz[k] = acc * g;
                }
#pragma unroll
                for (int k = 0; k < H; ++k) dz[k] = nz[k];
            }
        }
        __syncthreads();                                  // the next tile overwrites IN / Z / D / HP
    }

    // ---- this CTA's gradient sums -> global gradient (one atomic per weight per CTA) ---------------------------------
    if (kRegAcc) {
#pragma unroll
        for (int q = 0; q < kMaxL; ++q)
            if (q < L) flush_outer<ROWS, SLOTS>(H, (q == 0) ? in_dim : H, gacc[q], P.grad + ((q == 0) ? 0 : in_dim * H + (q - 1) * H * H));
    }
    flush_outer<ROWS, 1>(2, H, gout, P.grad + in_dim * H + (L - 1) * H * H);

    finish_step<ROWS>(P, n_w, loss_acc * inv_n);
}

// ------------------------------------------------------------------------------------------------
// 32-wide nets, TWO threads per row.  Every row keeps 188 floats of pre-activations in shared memory, so an SM holds
// ~256 rows whatever the tile shape: with one thread per row that is 8 warps per SM (2 per scheduler), and the kernel
// above waits on its own shared-memory loads (ncu: issue slots 43 % busy, short-scoreboard stalls).  Here thread
// (row, half) owns 16 of the 32 neurons of every layer -- half the FMAs of the forward / backward products per thread,
// the same shared-memory footprint, twice the warps.  The halves of a row sit in different warps (half = tid / ROWS is
// warp-uniform), meet through the shared-memory columns (one __syncthreads per layer) and recompute silu(z) of the row's
// layer input redundantly in the forward product.  Sums run in the same order as in the one-thread kernel except the
// two-row output layer (two 16-term partial sums).
// ------------------------------------------------------------------------------------------------
template <int ROWS, int RS, int THREADS, int SLOTS>
__device__ __forceinline__ void tile_outer_t(const float* __restrict__ D, int J, const float* __restrict__ P, int K,
                                             float (*acc)[4]) {
    const int kb = (K + 3) >> 2;
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) {        // work item o = tid + sl * THREADS of the J * kb (<= 256) of a layer
        const int o = threadIdx.x + sl * THREADS;
        if (o < J * kb) tile_outer_one<ROWS, RS>(D, P, K, o, acc[sl]);
    }
}

template <int ROWS>
__global__ void __launch_bounds__(2 * ROWS, 2) flow_matching_step_pair_kernel(const TrainParams P) {
    constexpr int H = 32, HH = 16, THREADS = 2 * ROWS;
    extern __shared__ __align__(16) float smem[];
    const int in_dim = P.in_dim, L = P.n_hidden;
    const int n_w = in_dim * H + (L - 1) * H * H + 2 * H;
    constexpr int RS = TrainSmem<H, ROWS>::RS;
    constexpr int kMaxL = 4;
    float* Wt = smem;                                    // transposed image: W1t [in][H], W2t.. [H][H], Woutt [H][2]
    float* col = smem + ((n_w + 3) & ~3);
    float* Z = col;                                      // [L][H][RS]
    float* IN = Z + (size_t)L * H * RS;                  // [kIn][RS]
    float* D = IN + (size_t)TrainSmem<H, ROWS>::kIn * RS;   // [H][RS]  delta_z of the current layer (rows 2..5: output-layer partial sums)
    float* HP = D + (size_t)H * RS;                      // [H][RS]  input of the current layer (h_{l-1})
    const int tid = threadIdx.x;
    const int half = tid / ROWS, row = tid - half * ROWS, jb = half * HH;
    constexpr int SLOTS = (H * (H / 4) + THREADS - 1) / THREADS;      // 1 (256 threads) or 2 (192 threads: 96-row tiles)
    float gacc[kMaxL][SLOTS][4], gout[1][4];             // this thread's work items of dW_1..dW_L and dWout, summed over the CTA's tiles
#pragma unroll
    for (int l = 0; l < kMaxL; ++l)
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl) gacc[l][sl][0] = gacc[l][sl][1] = gacc[l][sl][2] = gacc[l][sl][3] = 0.0f;
    gout[0][0] = gout[0][1] = gout[0][2] = gout[0][3] = 0.0f;
    {   // stage the weights, transposed (torch layout [out][in] -> [in][out])
        const float* w = P.weights;
        float* t = Wt;
        for (int l = 0; l <= L; ++l) {
            const int R = (l == L) ? 2 : H, C = (l == 0) ? in_dim : H;
            for (int i = tid; i < R * C; i += THREADS) { const int j = i / C, k = i - j * C; t[k * R + j] = w[i]; }
            w += R * C; t += R * C;
        }
    }
    __syncthreads();

    const long long n = P.n;
    const long long n_tiles = (n + ROWS - 1) / ROWS;
    const float inv_nm1 = (n > 1) ? 1.0f / (float)(n - 1) : 0.0f;
    const float inv_n = 1.0f / (float)n;
    float loss_acc = 0.0f;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long i = tile * ROWS + row;
        const bool valid = i < n;
        const long long ic = valid ? i : n - 1;
        // ---- inputs: half 0 the state columns and the regression target, half 1 PE5(wi) -----------------------------------
        float t0 = 0.0f, t1 = 0.0f;
        const int k0 = (P.domain == kDisk) ? 3 : 4;
        if (half == 0) {
            const float2 a = reinterpret_cast<const float2*>(P.x0)[ic];
            float2 b = reinterpret_cast<const float2*>(P.x1)[ic];
            const float alpha = P.alpha ? P.alpha[ic] : (float)ic * inv_nm1;
            t0 = b.x - a.x; t1 = b.y - a.y;
            if (P.domain == kDisk) {
                IN[0 * RS + row] = (1.0f - alpha) * a.x + alpha * b.x;
                IN[1 * RS + row] = (1.0f - alpha) * a.y + alpha * b.y;
                IN[2 * RS + row] = alpha;
            } else {
                const float kPi = 3.14159265358979323846f, kTwoPi = 6.28318530717958647692f;
                if (t1 < -kPi) { b.y += kTwoPi; t1 += kTwoPi; } else if (t1 > kPi) { b.y -= kTwoPi; t1 -= kTwoPi; }
                const float th = (1.0f - alpha) * a.x + alpha * b.x, ph = (1.0f - alpha) * a.y + alpha * b.y;
                float sn, cs;
                sincosf(ph, &sn, &cs);
                IN[0 * RS + row] = th; IN[1 * RS + row] = sn; IN[2 * RS + row] = cs; IN[3 * RS + row] = alpha;
            }
        } else {
            const float2 wi = reinterpret_cast<const float2*>(P.wi)[ic];
            float e[kPE5];
            positional_encoding<5>(wi.x, wi.y, e);
#pragma unroll
            for (int k = 0; k < kPE5; ++k) IN[(k0 + k) * RS + row] = e[k];
        }
        __syncthreads();

        // ---- forward: this thread's 16 neurons of every layer ------------------------------------------------------------
        // h_l = silu(z_l) is written next to z_l by the thread that owns the neuron (HP and D take turns as the layer-input
        // buffer; the last layer lands in HP, which the backward pass expects), so the product loop below is loads + FMAs only
        const float* Wl = Wt;
        for (int l = 0; l < L; ++l) {
            const int K = (l == 0) ? in_dim : H;
            float az[HH];
#pragma unroll
            for (int j = 0; j < HH; ++j) az[j] = 0.0f;
            const float* hin = (l == 0) ? IN : (((L - l) & 1) ? D : HP);
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                const float hk = hin[k * RS + row];
                const float4* w4 = reinterpret_cast<const float4*>(Wl + k * H + jb);
#pragma unroll
                for (int j4 = 0; j4 < HH / 4; ++j4) {
                    const float4 w = w4[j4];
                    az[4 * j4 + 0] = fmaf(hk, w.x, az[4 * j4 + 0]);
                    az[4 * j4 + 1] = fmaf(hk, w.y, az[4 * j4 + 1]);
                    az[4 * j4 + 2] = fmaf(hk, w.z, az[4 * j4 + 2]);
                    az[4 * j4 + 3] = fmaf(hk, w.w, az[4 * j4 + 3]);
                }
            }
            float* zo = Z + (size_t)l * H * RS;
            float* ho = ((L - 1 - l) & 1) ? D : HP;
#pragma unroll
            for (int j = 0; j < HH; ++j) {
                zo[(jb + j) * RS + row] = az[j];
                ho[(jb + j) * RS + row] = az[j] * sigmoid_newton(az[j]);
            }
            Wl += K * H;
            __syncthreads();                              // the row's other half of z_l / h_l
        }
        // ---- output layer + loss: each half sums its 16 neurons of h_L (in HP); D rows 0..1 <- delta_out -----------------
        const float* zl = Z + (size_t)(L - 1) * H * RS;
        {
            float q0 = 0.0f, q1 = 0.0f;
#pragma unroll 4
            for (int kk = 0; kk < HH; ++kk) {
                const int k = jb + kk;
                const float h = HP[k * RS + row];
                const float2 w = reinterpret_cast<const float2*>(Wl)[k];
                q0 = fmaf(h, w.x, q0); q1 = fmaf(h, w.y, q1);
            }
            D[(2 + 2 * half) * RS + row] = q0; D[(3 + 2 * half) * RS + row] = q1;
        }
        __syncthreads();
        const float p0 = D[2 * RS + row] + D[4 * RS + row], p1 = D[3 * RS + row] + D[5 * RS + row];
        // half 1 does not know the target: half 0 publishes delta_out, both read it back
        if (half == 0) {
            float d0 = valid ? (p0 - t0) : 0.0f, d1 = valid ? (p1 - t1) : 0.0f;
            loss_acc += 0.5f * (d0 * d0 + d1 * d1);
            D[0 * RS + row] = d0 * inv_n; D[1 * RS + row] = d1 * inv_n;
        }
        __syncthreads();
        const float d0 = D[0 * RS + row], d1 = D[1 * RS + row];

        // ---- backward ----------------------------------------------------------------------------------------------------
        tile_outer_t<ROWS, RS, THREADS, 1>(D, 2, HP, H, gout);              // dWout [2][H]
        float dz[HH];                                                         // delta_z of the last hidden layer, my 16 neurons
#pragma unroll
        for (int kk = 0; kk < HH; ++kk) {
            const int k = jb + kk;
            const float2 w = reinterpret_cast<const float2*>(Wl)[k];
            float h, g;
            silu_both(zl[k * RS + row], h, g);
            dz[kk] = (d0 * w.x + d1 * w.y) * g;
        }
        for (int l = L - 1; l >= 0; --l) {
            const int K = (l == 0) ? in_dim : H;
            const int off = (l == 0) ? 0 : in_dim * H + (l - 1) * H * H;
            __syncthreads();                              // everyone is done reading D / HP of the layer above
#pragma unroll
            for (int j = 0; j < HH; ++j) D[(jb + j) * RS + row] = dz[j];
            const float* src = (l == 0) ? IN : Z + (size_t)(l - 1) * H * RS;
            if (l > 0) {
#pragma unroll 4
                for (int kk = 0; kk < HH; ++kk) { const float z = src[(jb + kk) * RS + row]; HP[(jb + kk) * RS + row] = z * sigmoid_newton(z); }
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < kMaxL; ++q)               // the layer index selects the register block at compile time
                if (q == l) tile_outer_t<ROWS, RS, THREADS, SLOTS>(D, H, (l == 0) ? IN : HP, K, gacc[q]);
            if (l > 0) {
                // delta_h_{l-1}[k] = sum_j Wt_l[k][j] delta_z_l[j] over ALL 32 j (the other half's come from D), for my 16 k
                const float* Wc = Wt + off;
                float dzf[H];
#pragma unroll
                for (int j = 0; j < H; ++j) dzf[j] = D[j * RS + row];
                float nz[HH];
#pragma unroll 4
                for (int kk = 0; kk < HH; ++kk) {
                    const int k = jb + kk;
                    const float4* w4 = reinterpret_cast<const float4*>(Wc + k * H);
                    float acc = 0.0f;
#pragma unroll
                    for (int j4 = 0; j4 < H / 4; ++j4) {
                        const float4 w = w4[j4];
                        acc = fmaf(w.x, dzf[4 * j4 + 0], acc); acc = fmaf(w.y, dzf[4 * j4 + 1], acc);
                        acc = fmaf(w.z, dzf[4 * j4 + 2], acc); acc = fmaf(w.w, dzf[4 * j4 + 3], acc);
                    }
                    float h, g;
                    silu_both(src[k * RS + row], h, g);
                    nz[kk] = acc * g;
                }
#pragma unroll
                for (int kk = 0; kk < HH; ++kk) dz[kk] = nz[kk];
            }
        }
        __syncthreads();                                  // the next tile overwrites IN / Z / D / HP
    }

    // ---- this CTA's gradient sums -> global gradient (one atomic per weight per CTA) ---------------------------------
#pragma unroll
    for (int q = 0; q < kMaxL; ++q)
        if (q < L) {
            const int K = (q == 0) ? in_dim : H, kb = (K + 3) >> 2;
#pragma unroll
            for (int sl = 0; sl < SLOTS; ++sl)
                if (tid + sl * THREADS < H * kb)
                    add_outer_one(K, tid + sl * THREADS, gacc[q][sl], P.grad + ((q == 0) ? 0 : in_dim * H + (q - 1) * H * H));
        }
    if (tid < 2 * (H / 4)) add_outer_one(H, tid, gout[0], P.grad + in_dim * H + (L - 1) * H * H);
    finish_step<THREADS>(P, n_w, loss_acc * inv_n);
}

// ------------------------------------------------------------------------------------------------
// pretrain stage: negative log-likelihood of the base distribution (disk_domain_sampling.py:14-33,
// spherical_domain_sampling.py:16-35):  loss = -mean(D_base.log_prob(omega_o, omega_i));  Adam.
// Base net 14 -> 16 -> 4 with biases (model.py:374-398 disk, :277-317 spherical); parameters in the 308-float base-blob
// order W1 [16,14], b1 [16], Wo [4,16], bo [4].  Same structure as the flow kernel: thread <-> row forward / backward,
// the parameter gradients as tile products [16 x rows] . [rows x 15] (PE3 + a ones column for the bias) and
// [4 x rows] . [rows x 17], summed in registers over the CTA's tiles.
// ------------------------------------------------------------------------------------------------
// d/dx log I0(x) with the polynomials of torch.distributions.von_mises._log_modified_bessel_fn(order=0)
__device__ __forceinline__ float dlog_i0(float x) {
    if (x < 3.75f) {
        const float y = (x / 3.75f) * (x / 3.75f);
        const float c[7] = {1.0f, 3.5156229f, 3.0899424f, 1.2067492f, 0.2659732f, 0.360768e-1f, 0.45813e-2f};
        float r = c[6], dr = 0.0f;
#pragma unroll
        for (int i = 5; i >= 0; --i) { dr = fmaf(dr, y, r); r = fmaf(r, y, c[i]); }
        return dr / r * (2.0f * x / (3.75f * 3.75f));
    }
    const float y = 3.75f / x;
    const float c[9] = {0.39894228f, 0.1328592e-1f, 0.225319e-2f, -0.157565e-2f, 0.916281e-2f, -0.2057706e-1f,
                        0.2635537e-1f, -0.1647633e-1f, 0.392377e-2f};
    float r = c[8], dr = 0.0f;
#pragma unroll
    for (int i = 7; i >= 0; --i) { dr = fmaf(dr, y, r); r = fmaf(r, y, c[i]); }
    return 1.0f - 0.5f / x + dr / r * (-3.75f / (x * x));
}

template <int ROWS>
__global__ void __launch_bounds__(ROWS) base_nll_step_kernel(const TrainParams P) {
    constexpr int RS = ROWS + 4;
    __shared__ __align__(16) float Wb[kBaseFloats + 4];
    __shared__ __align__(16) float E[15 * RS], Hc[17 * RS], DZ[16 * RS], DP[4 * RS];
    const int tid = threadIdx.x;
    for (int i = tid; i < kBaseFloats; i += ROWS) Wb[i] = P.weights[i];
    __syncthreads();
    const long long n = P.n, n_tiles = (n + ROWS - 1) / ROWS;
    const float inv_n = 1.0f / (float)n;
    float loss_acc = 0.0f;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};               // work item tid: < 64 -> dW1/db1 item, 64..83 -> dWo/dbo item
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long i = tile * ROWS + tid;
        const bool valid = i < n;
        const long long ic = valid ? i : n - 1;
        const float2 x = reinterpret_cast<const float2*>(P.x1)[ic];
        const float2 wi = reinterpret_cast<const float2*>(P.wi)[ic];
        float e[kPE3];
        positional_encoding<3>(wi.x, wi.y, e);
        float z[16], pr[4] = {Wb[304], Wb[305], Wb[306], Wb[307]};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float a = Wb[224 + j];
#pragma unroll
            for (int k = 0; k < kPE3; ++k) a = fmaf(e[k], Wb[j * kPE3 + k], a);
            z[j] = a;
            const float h = a * sigmoid_precise(a);
            Hc[j * RS + tid] = h;
            pr[0] = fmaf(h, Wb[240 + j], pr[0]); pr[1] = fmaf(h, Wb[256 + j], pr[1]);
            pr[2] = fmaf(h, Wb[272 + j], pr[2]); pr[3] = fmaf(h, Wb[288 + j], pr[3]);
        }
        Hc[16 * RS + tid] = 1.0f;
#pragma unroll
        for (int k = 0; k < kPE3; ++k) E[k * RS + tid] = e[k];
        E[14 * RS + tid] = 1.0f;
        // log-density and its gradient w.r.t. the four outputs
        float lp, g[4];
        if (P.domain == kDisk) {
            const float s0 = expf(pr[2]), s1 = expf(pr[3]);
            const float e0 = (x.x - pr[0]) / s0, e1 = (x.y - pr[1]) / s1;
            lp = -kLog2Pi - (pr[2] + pr[3]) - 0.5f * (e0 * e0 + e1 * e1);
            g[0] = e0 / s0; g[1] = e1 / s1; g[2] = e0 * e0 - 1.0f; g[3] = e1 * e1 - 1.0f;
        } else {
            const float ex = expf(pr[1]), sc = ex + 1e-3f;
            const float ee = (x.x - pr[0]) / sc;
            const float kappa = softplus_torch(pr[3]) + 1e-3f;
            float sn, cs;
            sincosf(x.y - pr[2], &sn, &cs);
            lp = -0.5f * kLog2Pi - pr[1] - 0.5f * ee * ee + kappa * cs - kLog2Pi - log_i0(kappa);
            g[0] = ee / sc;
            g[1] = ee * ee * ex / sc - 1.0f;
            g[2] = kappa * sn;
            g[3] = (cs - dlog_i0(kappa)) * (pr[3] > 20.0f ? 1.0f : sigmoid_precise(pr[3]));
        }
        const float wgt = valid ? -inv_n : 0.0f;          // loss = -mean(logp)
        if (valid) loss_acc -= lp;
        float dh[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) dh[j] = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float d = g[c] * wgt;
            DP[c * RS + tid] = d;
#pragma unroll
            for (int j = 0; j < 16; ++j) dh[j] = fmaf(Wb[240 + c * 16 + j], d, dh[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float sg = sigmoid_precise(z[j]);
            DZ[j * RS + tid] = dh[j] * sg * fmaf(z[j], 1.0f - sg, 1.0f);
        }
        __syncthreads();
        if (tid < 64) tile_outer_one<ROWS, RS>(DZ, E, 15, tid, acc);                // [16] x [15]: kb = 4 -> 64 items
        else if (tid < 84) tile_outer_one<ROWS, RS>(DP, Hc, 17, tid - 64, acc);     // [4] x [17]:  kb = 5 -> 20 items
        __syncthreads();
    }
    // this CTA's sums -> global gradient in blob order
    if (tid < 64) {
        const int j = tid / 4, kq = tid % 4;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int k = kq + 4 * b;
            if (k < 14) atomicAdd(P.grad + j * kPE3 + k, acc[b]);
            else if (k == 14) atomicAdd(P.grad + 224 + j, acc[b]);
        }
    } else if (tid < 84) {
        const int o = tid - 64, c = o / 5, kq = o % 5;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int k = kq + 5 * b;
            if (k < 16) atomicAdd(P.grad + 240 + c * 16 + k, acc[b]);
            else if (k == 16) atomicAdd(P.grad + 304 + c, acc[b]);
        }
    }
    finish_step<ROWS>(P, kBaseFloats, loss_acc * inv_n);
}

int launch_base_nll_step(const TrainParams& P, cudaStream_t stream) {
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto kern = base_nll_step_kernel<128>;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, 0) != cudaSuccess || occ < 1) return -3;
    long long tiles = (P.n + 127) / 128, grid = (long long)sms * occ;
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, 128, 0, stream>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

template <int H, int ROWS>
static int launch_train_t(const TrainParams& P, cudaStream_t stream) {
    const size_t smem = sizeof(float) * TrainSmem<H, ROWS>::floats(P.in_dim, P.n_hidden) + 16;
    if (smem > 227u * 1024u) return -2;
    auto kern = flow_matching_step_kernel<H, ROWS>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, ROWS, smem) != cudaSuccess || occ < 1) return -3;
    long long tiles = (P.n + ROWS - 1) / ROWS, grid = (long long)sms * occ;
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, ROWS, smem, stream>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

template <int ROWS>
static int launch_train_pair_t(const TrainParams& P, cudaStream_t stream) {
    const size_t smem = sizeof(float) * TrainSmem<32, ROWS>::floats(P.in_dim, P.n_hidden) + 16;
    if (smem > 227u * 1024u) return -2;
    auto kern = flow_matching_step_pair_kernel<ROWS>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 2 * ROWS, smem) != cudaSuccess || occ < 1) return -3;
    long long tiles = (P.n + ROWS - 1) / ROWS, grid = (long long)sms * occ;
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, 2 * ROWS, smem, stream>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

int launch_flow_matching_step(const TrainParams& P, cudaStream_t stream) {
    if (P.n_hidden < 1) return -2;
    // 32-wide nets with up to 4 hidden layers: two threads per row (BSDFDIFF_TRAIN_PAIR=0 selects the one-thread kernel for A/B runs)
    static const bool pair = [] { const char* e = getenv("BSDFDIFF_TRAIN_PAIR"); return !(e && e[0] == '0'); }();
    if (pair && P.hidden == 32 && P.n_hidden <= 4)
        return P.n_hidden <= 3 ? launch_train_pair_t<128>(P, stream) : launch_train_pair_t<96>(P, stream);
    // 32-wide nets: 128-row tiles while two CTAs fit an SM (3 hidden layers: 111 KB each); with 4 hidden layers a 128-row
    // CTA needs 133 KB (one per SM), 96-row tiles need 104 KB (two per SM: 6 warps instead of 4)
    if (P.hidden == 32) return P.n_hidden <= 3 ? launch_train_t<32, 128>(P, stream)
                             : P.n_hidden == 4 ? launch_train_t<32, 96>(P, stream) : -2;
    if (P.hidden == 64 && P.n_hidden > 6) return -2;
    // 64-wide nets: the z columns of 6 layers + 89 KB of weights only fit with 32-row tiles
    if (P.hidden == 64) return P.n_hidden <= 4 ? launch_train_t<64, 64>(P, stream) : launch_train_t<64, 32>(P, stream);
    return -2;
}

}  // namespace bsdfdiff
