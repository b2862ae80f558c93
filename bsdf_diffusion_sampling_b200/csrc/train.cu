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

namespace bsdfdiff {

__device__ __forceinline__ void silu_both(float z, float& h, float& g) {
    const float s = 1.0f / (1.0f + expf(-z));
    h = z * s;
    g = s * fmaf(z, 1.0f - s, 1.0f);
}

template <int H, int ROWS>
struct TrainSmem {
    // per-thread columns, element e of row r at col[e * ROWS + r]
    static constexpr int kIn = 28;                       // first-layer input (25 | 26), padded
    __host__ __device__ static size_t floats(int in_dim, int n_hidden) {
        const size_t w = (size_t)in_dim * H + (size_t)(n_hidden - 1) * H * H + 2 * H;
        return ((w + 3) & ~(size_t)3) + (size_t)ROWS * ((size_t)n_hidden * H + kIn + 2 * H);
    }
};

// dW[j][k] (+)= sum_r D[j][r] * P[k][r] over the tile; outputs (j, k) are dealt to the CTA's threads in blocks of 4 k's.
// D: [J][ROWS], P: [K][ROWS] (shared-memory columns); grad row-major [J][K] (torch layout)
template <int ROWS>
__device__ __forceinline__ void tile_outer(const float* __restrict__ D, int J, const float* __restrict__ P, int K,
                                           float* __restrict__ grad) {
    const int kb = (K + 3) >> 2;                         // blocks of 4 consecutive k
    for (int o = threadIdx.x; o < J * kb; o += ROWS) {
        const int j = o / kb, k0 = (o - j * kb) << 2;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float4* d4 = reinterpret_cast<const float4*>(D + (size_t)j * ROWS);
        const float4* p0 = reinterpret_cast<const float4*>(P + (size_t)k0 * ROWS);
        const float4* p1 = reinterpret_cast<const float4*>(P + (size_t)min(k0 + 1, K - 1) * ROWS);
        const float4* p2 = reinterpret_cast<const float4*>(P + (size_t)min(k0 + 2, K - 1) * ROWS);
        const float4* p3 = reinterpret_cast<const float4*>(P + (size_t)min(k0 + 3, K - 1) * ROWS);
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
        float* g = grad + (size_t)j * K + k0;
        atomicAdd(g, a0);
        if (k0 + 1 < K) atomicAdd(g + 1, a1);
        if (k0 + 2 < K) atomicAdd(g + 2, a2);
        if (k0 + 3 < K) atomicAdd(g + 3, a3);
    }
}

template <int H, int ROWS>
__global__ void __launch_bounds__(ROWS) flow_matching_step_kernel(const TrainParams P) {
    extern __shared__ __align__(16) float smem[];
    const int in_dim = P.in_dim, L = P.n_hidden;
    const int n_w = in_dim * H + (L - 1) * H * H + 2 * H;
    float* Wt = smem;                                    // transposed image: W1t [in][H], W2t.. [H][H], Woutt [H][2]
    float* col = smem + ((n_w + 3) & ~3);
    float* Z = col;                                      // [L][H][ROWS]
    float* IN = Z + (size_t)L * H * ROWS;                // [kIn][ROWS]
    float* D = IN + (size_t)TrainSmem<H, ROWS>::kIn * ROWS;   // [H][ROWS]  delta_z of the current layer
    float* HP = D + (size_t)H * ROWS;                    // [H][ROWS]  input of the current layer (h_{l-1})
    const int tid = threadIdx.x;

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
            IN[0 * ROWS + tid] = (1.0f - alpha) * a.x + alpha * b.x;
            IN[1 * ROWS + tid] = (1.0f - alpha) * a.y + alpha * b.y;
            IN[2 * ROWS + tid] = alpha;
            k0 = 3;
        } else {
            // spherical_domain_sampling.py:61-72: move omega_o.phi to within pi of x_0.phi, interpolate, embed periodically
            const float kPi = 3.14159265358979323846f, kTwoPi = 6.28318530717958647692f;
            if (t1 < -kPi) { b.y += kTwoPi; t1 += kTwoPi; } else if (t1 > kPi) { b.y -= kTwoPi; t1 -= kTwoPi; }
            const float th = (1.0f - alpha) * a.x + alpha * b.x, ph = (1.0f - alpha) * a.y + alpha * b.y;
            float s, c;
            sincosf(ph, &s, &c);
            IN[0 * ROWS + tid] = th; IN[1 * ROWS + tid] = s; IN[2 * ROWS + tid] = c; IN[3 * ROWS + tid] = alpha;
            k0 = 4;
        }
        {
            float e[kPE5];
            positional_encoding<5>(wi.x, wi.y, e);
#pragma unroll
            for (int k = 0; k < kPE5; ++k) IN[(k0 + k) * ROWS + tid] = e[k];
        }

        // ---- forward ---------------------------------------------------------------------------------------
        const float* Wl = Wt;
        for (int l = 0; l < L; ++l) {
            const int K = (l == 0) ? in_dim : H;
            float az[H];
#pragma unroll
            for (int j = 0; j < H; ++j) az[j] = 0.0f;
            const float* zin = Z + (size_t)(l - 1) * H * ROWS;
            for (int k = 0; k < K; ++k) {
                float hk;
                if (l == 0) hk = IN[k * ROWS + tid];
                else { const float z = zin[k * ROWS + tid]; hk = z / (1.0f + expf(-z)); }
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
            float* zo = Z + (size_t)l * H * ROWS;
#pragma unroll
            for (int j = 0; j < H; ++j) zo[j * ROWS + tid] = az[j];
            Wl += K * H;
        }
        // output layer + loss; HP <- h_L, D rows 0..1 <- delta_out
        float p0 = 0.0f, p1 = 0.0f;
        {
            const float* zl = Z + (size_t)(L - 1) * H * ROWS;
            for (int k = 0; k < H; ++k) {
                const float z = zl[k * ROWS + tid];
                const float h = z / (1.0f + expf(-z));
                HP[k * ROWS + tid] = h;
                const float2 w = reinterpret_cast<const float2*>(Wl)[k];
                p0 = fmaf(h, w.x, p0); p1 = fmaf(h, w.y, p1);
            }
        }
        float d0 = valid ? (p0 - t0) : 0.0f, d1 = valid ? (p1 - t1) : 0.0f;
        loss_acc += 0.5f * (d0 * d0 + d1 * d1);
        d0 *= inv_n; d1 *= inv_n;
        D[0 * ROWS + tid] = d0; D[1 * ROWS + tid] = d1;
        __syncthreads();

        // ---- backward ----------------------------------------------------------------------------------------
        const int off_out = in_dim * H + (L - 1) * H * H;
        tile_outer<ROWS>(D, 2, HP, H, P.grad + off_out);                  // dWout [2][H]
        // delta_z of the last hidden layer
        float dz[H];
        {
            const float* zl = Z + (size_t)(L - 1) * H * ROWS;
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const float2 w = reinterpret_cast<const float2*>(Wl)[k];
                float h, g;
                silu_both(zl[k * ROWS + tid], h, g);
                dz[k] = (d0 * w.x + d1 * w.y) * g;
            }
        }
        for (int l = L - 1; l >= 0; --l) {
            const int K = (l == 0) ? in_dim : H;
            const int off = (l == 0) ? 0 : in_dim * H + (l - 1) * H * H;
            __syncthreads();                              // everyone is done reading D / HP of the layer above
#pragma unroll
            for (int j = 0; j < H; ++j) D[j * ROWS + tid] = dz[j];
            const float* src = (l == 0) ? IN : Z + (size_t)(l - 1) * H * ROWS;
            if (l > 0) {
                for (int k = 0; k < H; ++k) { const float z = src[k * ROWS + tid]; HP[k * ROWS + tid] = z / (1.0f + expf(-z)); }
            }
            __syncthreads();
            tile_outer<ROWS>(D, H, (l == 0) ? IN : HP, K, P.grad + off);   // dW_l [H][K]
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
                    silu_both(src[k * ROWS + tid], h, g);
                    nz[k] = acc * g;
                }
#pragma unroll
                for (int k = 0; k < H; ++k) dz[k] = nz[k];
            }
        }
        __syncthreads();                                  // the next tile overwrites IN / Z / D / HP
    }

    // ---- loss, then Adam by the last CTA -------------------------------------------------------------------------
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
    if ((tid & 31) == 0 && loss_acc != 0.0f) atomicAdd(P.loss, loss_acc * inv_n);      // mean over [N,2]: sum / (2N), 0.5 folded above
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
        for (int i = tid; i < n_w; i += ROWS) {
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

int launch_flow_matching_step(const TrainParams& P, cudaStream_t stream) {
    if (P.n_hidden < 1 || P.n_hidden > 8) return -2;
    if (P.hidden == 32) return launch_train_t<32, 128>(P, stream);
    if (P.hidden == 64) return launch_train_t<64, 64>(P, stream);     // 64-row tiles: the z columns of 6 layers + 89 KB of weights fill the SM
    return -2;
}

}  // namespace bsdfdiff
