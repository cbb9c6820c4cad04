// policy_kernels.cuh — the rollout policy on the device: MADDPG.step() of the reference
// (marl_llm/algorithm/algorithms/maddpg.py:72-87 -> utils/agents.py:69-96 -> utils/networks.py:33-44) evaluated for every
// agent of every env straight from the simulator's observation layout.
//
//   obs [E][K0][n_a] f32 (feature-major, what k_step writes)  ->  act [E][A][n_a] f32 (what k_step reads)
//   h1 = leaky_relu(W1 obs + b1); h2 = leaky_relu(W2 h1 + b2); h3 = leaky_relu(W3 h2 + b3); a = tanh(W4 h3 + b4)
//   explore (agents.py:85-93): a += scale * N(0,1), clamp to [-1,1]  |  or a = U(-1,1) for the whole batch (epsilon branch)
//
// k_policy_mlp is the exact-fp32 path: plain FFMA, fp32 accumulation in a fixed k-ascending order (deterministic; agrees
// with torch's fp32 Linear to rounding).  A "column" is one agent; a CTA owns POL_M columns, keeps their activations in
// shared memory as [k][column] for all four layers (nothing but obs is read from and nothing but act written to HBM),
// and streams the (L2-resident, pre-transposed, zero-padded) weights through a double-buffered cp.async ring.
// Each thread accumulates a 6 (outputs) x 8 (columns) register tile: 48 FFMA per 3 + 2 shared-memory vector loads.
#pragma once
#include <cuda_runtime.h>

#include <stdint.h>

namespace swarm {

constexpr int POL_M = 64;        // columns (agents) per CTA
constexpr int POL_HP = 192;      // padded hidden width and padded input width (reference: hidden 180, obs 192)
constexpr int POL_KC = 16;       // k rows per weight chunk
constexpr int POL_THREADS = 256; // 32 output tiles of 6  x  8 column tiles of 8
constexpr int POL_AMAX = 8;      // largest action dimension

struct PolicyParams {
    const float *obs; float *act; float *log_pi;        // log_pi [E][1][n_a] or NULL
    long n_cols;                 // E * n_a
    int n_a, K0, A;              // agents per env, observation features (<= POL_HP), action dimension (<= POL_AMAX)
    const float *Wt[3];          // [POL_HP][POL_HP] transposed weights Wt[k][n] of the three hidden layers, zero-padded
    const float *b[3];           // [POL_HP] biases, zero-padded
    const float *W4;             // [A][POL_HP] output layer (row-major like torch), zero-padded
    const float *b4;             // [A]
    float slope;                 // leaky_relu negative slope (torch default 0.01, networks.py:11)
    int explore;                 // 0 none | 1 gaussian | 2 uniform (agents.py:85-93)
    float scale;                 // noise scale (utils/noise.py:30)
    uint64_t seed, step;         // keys of the counter-based generator
};

__device__ __forceinline__ uint64_t pol_mix64(uint64_t seed, uint64_t a, uint64_t b, uint64_t k) {
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + a * 0xBF58476D1CE4E5B9ull + b * 0x94D049BB133111EBull + k * 0xD6E8FEB86659FD93ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(POL_THREADS, 1) k_policy_mlp(const PolicyParams P) {
    extern __shared__ __align__(16) float psm[];
    float *hA = psm;                               // [POL_HP][POL_M] activations (ping)
    float *hB = hA + POL_HP * POL_M;               // [POL_HP][POL_M] activations (pong)
    float *wbuf = hB + POL_HP * POL_M;             // [2][POL_KC][POL_HP] weight chunks
    const int tid = threadIdx.x;
    const long c0 = (long)blockIdx.x * POL_M;      // first column of this CTA
    const int n_a = P.n_a;

    // ---- observation tile: hA[k][c] = obs[e][k][a], column c0 + c = e * n_a + a; rows K0..POL_HP-1 and dead columns = 0
    for (int idx = tid; idx < POL_HP * POL_M; idx += POL_THREADS) {
        const int k = idx / POL_M, c = idx - k * POL_M;
        const long col = c0 + c;
        float v = 0.f;
        if (k < P.K0 && col < P.n_cols) {
            const long e = col / n_a; const int a = (int)(col - e * n_a);
            v = P.obs[(e * P.K0 + k) * n_a + a];
        }
        hA[idx] = v;
    }

    // this thread's 6 x 8 output tile: outputs n0..n0+5, columns a0..a0+3 and a0+32..a0+35 (the 8 lanes sharing n0 read one
    // contiguous 128-byte line per vector load: conflict-free)
    const int n0 = (tid >> 3) * 6, a0 = (tid & 7) * 4;
    float *hin = hA, *hout = hB;
#pragma unroll 1
    for (int layer = 0; layer < 3; ++layer) {
        const float *Wt = P.Wt[layer];
        float acc[6][8];
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
        // chunk 0 in flight; the __syncthreads below also covers the activation tile written above / by the last layer
        for (int q = tid; q < POL_KC * POL_HP / 4; q += POL_THREADS) cp_async16(wbuf + q * 4, Wt + q * 4);
        cp_async_commit();
#pragma unroll 1
        for (int ch = 0; ch < POL_HP / POL_KC; ++ch) {
            if (ch + 1 < POL_HP / POL_KC) {
                float *dst = wbuf + ((ch + 1) & 1) * POL_KC * POL_HP;
                const float *src = Wt + (size_t)(ch + 1) * POL_KC * POL_HP;
                for (int q = tid; q < POL_KC * POL_HP / 4; q += POL_THREADS) cp_async16(dst + q * 4, src + q * 4);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();                                   // chunk ch (and the input activations) visible to all
            const float *w = wbuf + (ch & 1) * POL_KC * POL_HP;
#pragma unroll 4
            for (int kk = 0; kk < POL_KC; ++kk) {
                const int k = ch * POL_KC + kk;
                const float2 w01 = *reinterpret_cast<const float2 *>(w + kk * POL_HP + n0);
                const float2 w23 = *reinterpret_cast<const float2 *>(w + kk * POL_HP + n0 + 2);
                const float2 w45 = *reinterpret_cast<const float2 *>(w + kk * POL_HP + n0 + 4);
                const float4 x0 = *reinterpret_cast<const float4 *>(hin + k * POL_M + a0);
                const float4 x1 = *reinterpret_cast<const float4 *>(hin + k * POL_M + a0 + 32);
                const float wr[6] = {w01.x, w01.y, w23.x, w23.y, w45.x, w45.y};
                const float xr[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(wr[r], xr[c], acc[r][c]);
            }
            __syncthreads();                                   // everyone is done with buffer ch & 1 before it is refilled
        }
        const float *bias = P.b[layer];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const float bn = __ldg(bias + n0 + r);
            float o[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) { const float v = acc[r][c] + bn; o[c] = v > 0.f ? v : v * P.slope; }   // F.leaky_relu
            *reinterpret_cast<float4 *>(hout + (n0 + r) * POL_M + a0) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4 *>(hout + (n0 + r) * POL_M + a0 + 32) = make_float4(o[4], o[5], o[6], o[7]);
        }
        float *t = hin; hin = hout; hout = t;
    }
    __syncthreads();                                           // h3 complete

    // ---- output layer + tanh + exploration: one (column, action component) per thread
    for (int idx = tid; idx < POL_M * P.A; idx += POL_THREADS) {
        const int j = idx / POL_M, c = idx - j * POL_M;
        const long col = c0 + c;
        if (col >= P.n_cols) continue;
        const float *w4 = P.W4 + (size_t)j * POL_HP;
        float s = 0.f;
#pragma unroll 4
        for (int k = 0; k < POL_HP; ++k) s = fmaf(__ldg(w4 + k), hin[k * POL_M + c], s);
        float a = tanhf(s + __ldg(P.b4 + j));                  // networks.py:29,43 (constrain_out)
        float nz = 0.f;
        if (P.explore == 1) {                                  // agents.py:90-93: a += scale * N(0,1), clamp
            const uint64_t r = pol_mix64(P.seed, P.step, (uint64_t)col, (uint64_t)j);
            const float u1 = ((float)(r >> 40) + 1.0f) * (1.0f / 16777216.0f);          // (0, 1]
            const float u2 = (float)((r >> 16) & 0xffffffu) * (1.0f / 16777216.0f);     // [0, 1)
            nz = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2) * P.scale;                 // Box-Muller
            a = fminf(fmaxf(a + nz, -1.f), 1.f);
        } else if (P.explore == 2) {                           // agents.py:86-88: uniform action for the whole batch
            const uint64_t r = pol_mix64(P.seed, P.step, (uint64_t)col, (uint64_t)j);
            a = (float)(r >> 40) * (2.0f / 16777216.0f) - 1.0f;
        }
        const long e = col / n_a; const int ag = (int)(col - e * n_a);
        P.act[(e * P.A + j) * n_a + ag] = a;
        if (P.log_pi) {
            // agents.py:82,88,91 + noise.py:32-37; the components of one column live in lanes POL_M apart: recompute them
            float lp = 0.f;
            if (P.explore == 2) lp = -(float)P.A * 0.69314718056f;
            else if (P.explore == 1 && j == 0) {
                float q = 0.f;
                for (int jj = 0; jj < P.A; ++jj) {
                    const uint64_t r = pol_mix64(P.seed, P.step, (uint64_t)col, (uint64_t)jj);
                    const float u1 = ((float)(r >> 40) + 1.0f) * (1.0f / 16777216.0f);
                    const float u2 = (float)((r >> 16) & 0xffffffu) * (1.0f / 16777216.0f);
                    const float g = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);            // noise / scale
                    q += g * g;
                }
                lp = -0.5f * q - (float)P.A * logf(P.scale * 2.50662827463f);
            }
            if (j == 0) P.log_pi[e * n_a + ag] = lp;
        }
    }
}

}  // namespace swarm
