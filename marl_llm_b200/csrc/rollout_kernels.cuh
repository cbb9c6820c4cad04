// rollout_kernels.cuh — device-resident replay storage for the batched simulator.
//
// The reference keeps its replay buffer in host NumPy arrays (marl_llm/algorithm/utils/buffer_agent.py = BUF):
// push() transposes the env's feature-major arrays ([dim, n_a]) into agent rows and appends them to a ring
// (BUF:67-128), sample() gathers rows by index (BUF:130-177).  With the simulator on the GPU both are pure HBM
// data movement; these two kernels do them without leaving the device.
//
//   k_rollout_push    one CTA per (env, chunk of <= 32 agents): the [dim][n_a] tiles of obs and next_obs are read with
//                     coalesced row segments, transposed through shared memory and written as contiguous agent rows;
//                     the small per-agent arrays (action, prior, log_pi, reward, done) ride along.
//   k_rollout_gather  one warp per sampled row, 16-byte loads/stores.
// Storage is fp32: the reference stores fp64 and casts to fp32 in sample() (torch.Tensor(x), BUF:170-173); rounding
// once at push time yields the same fp32 values.
#pragma once
#include <cuda_runtime.h>

#include <stdint.h>

namespace swarm {

struct RolloutBuf {
    float *obs, *act, *act_prior, *log_pi, *rew, *next_obs, *done;   // [capacity][dim] row-major; act_prior / log_pi may be NULL
    int obs_dim, act_dim;
};

struct PushParams {
    RolloutBuf B;
    long row0;                 // first ring row this push writes (BUF:96-99 decided by the host mirror)
    int n_a, a0, a1;           // agents per env and the [a0, a1) slice stored (BUF:86-90 `index`)
    const void *obs, *next_obs, *rew, *prior;   // [E][dim][n_a] in the simulator's output dtype
    const void *act;           // [E][act_dim][n_a]
    const float *log_pi;       // [E][1][n_a] or NULL
    const unsigned char *done; // [E][1][n_a] (bool)
    int out_f32, act_f32;
    int parts;                 // bit 0: transpose obs, bit 1: transpose next_obs, bit 2: the small per-agent arrays
};

__device__ __forceinline__ float ldf(const void *p, int is_f32, size_t k) {
    return is_f32 ? reinterpret_cast<const float *>(p)[k] : (float)reinterpret_cast<const double *>(p)[k];
}

constexpr int PUSH_CHUNK = 32;      // agents per CTA
constexpr int PUSH_THREADS = 256;

__global__ void __launch_bounds__(PUSH_THREADS) k_rollout_push(const PushParams P) {
    extern __shared__ float tile[];                       // [obs_dim][PUSH_CHUNK + 1]
    const int e = blockIdx.x;
    const int span = P.a1 - P.a0;
    const int c0 = P.a0 + blockIdx.y * PUSH_CHUNK;        // first agent of this chunk
    const int nc = min(PUSH_CHUNK, P.a1 - c0);            // agents in it
    const int D = P.B.obs_dim, n_a = P.n_a;
    const long row = P.row0 + (long)e * span + (long)blockIdx.y * PUSH_CHUNK;   // ring row of agent c0 (BUF:86-90: rows follow agent order)
    for (int pass = 0; pass < 2; ++pass) {
        if (!(P.parts & (1 << pass))) continue;
        const void *src = pass ? P.next_obs : P.obs;
        float *dst = (pass ? P.B.next_obs : P.B.obs) + row * D;
        const size_t base = (size_t)e * D * n_a;
        // feature-major read: consecutive threads walk the agents of one feature row (coalesced; the whole tile is one
        // contiguous block when the chunk covers the env).  (d, a) advance incrementally: no division in the loops.
        {
            int d = threadIdx.x / nc, a = threadIdx.x - d * nc;
            const int sd = PUSH_THREADS / nc, sa = PUSH_THREADS - sd * nc;
#pragma unroll 4
            for (int k = threadIdx.x; k < D * nc; k += PUSH_THREADS) {
                tile[d * (PUSH_CHUNK + 1) + a] = ldf(src, P.out_f32, base + (size_t)d * n_a + c0 + a);
                d += sd; a += sa;
                if (a >= nc) { a -= nc; ++d; }
            }
        }
        __syncthreads();
        // agent-major write: the nc rows of this chunk are contiguous in the ring; the odd tile stride keeps the
        // transposed shared-memory reads conflict-free
        {
            int a = threadIdx.x / D, d = threadIdx.x - a * D;
            const int sa = PUSH_THREADS / D, sd = PUSH_THREADS - sa * D;
#pragma unroll 4
            for (int k = threadIdx.x; k < D * nc; k += PUSH_THREADS) {
                dst[k] = tile[d * (PUSH_CHUNK + 1) + a];
                a += sa; d += sd;
                if (d >= D) { d -= D; ++a; }
            }
        }
        __syncthreads();
    }
    if (!(P.parts & 4)) return;
    const int A = P.B.act_dim;
    for (int k = threadIdx.x; k < A * nc; k += PUSH_THREADS) {
        const int a = k / A, d = k - a * A;
        const size_t s = (size_t)e * A * n_a + (size_t)d * n_a + c0 + a;
        P.B.act[row * A + k] = ldf(P.act, P.act_f32, s);
        if (P.prior && P.B.act_prior) P.B.act_prior[row * A + k] = ldf(P.prior, P.out_f32, s);
    }
    for (int a = threadIdx.x; a < nc; a += PUSH_THREADS) {
        const size_t s = (size_t)e * n_a + c0 + a;
        P.B.rew[row + a] = ldf(P.rew, P.out_f32, s);
        P.B.done[row + a] = P.done[s] ? 1.0f : 0.0f;
        if (P.log_pi && P.B.log_pi) P.B.log_pi[row + a] = P.log_pi[s];
    }
}

// ---- TMA variant of the push (the 30-agent production shape): when one CTA covers a whole env and the tile is fp32,
// the [dim][n_a] tile is ONE contiguous block on both sides.  It is pulled into shared memory by a single bulk async copy
// (cp.async.bulk, UBLKCP), transposed shared -> shared by the threads, and pushed out by a single bulk store; the threads
// never touch global memory for the two big arrays.  blockIdx.y picks obs / next_obs.
__device__ __forceinline__ uint32_t rsmem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(PUSH_THREADS) k_rollout_push_tma(const PushParams P) {
    extern __shared__ __align__(128) float tile[];        // [D*n_a] in, [n_a*D] out
    __shared__ __align__(8) uint64_t bar;
    const int e = blockIdx.x;
    const int npass = (P.parts & 1) + ((P.parts >> 1) & 1);        // grid.y = max(npass, 1)
    const int pass = npass == 2 ? (int)blockIdx.y : (npass == 1 ? ((P.parts & 1) ? 0 : 1) : -1);
    const bool do_small = (P.parts & 4) && blockIdx.y == 0;
    const int D = P.B.obs_dim, n_a = P.n_a, n = D * n_a;
    const unsigned bytes = (unsigned)n * 4u;
    float *tin = tile, *tout = tile + n;
    const long row = P.row0 + (long)e * n_a;
    const float *src = reinterpret_cast<const float *>(pass == 1 ? P.next_obs : P.obs) + (size_t)e * n;
    float *dst = (pass == 1 ? P.B.next_obs : P.B.obs) + row * D;
    if (pass >= 0 && threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(rsmem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rsmem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(rsmem_u32(tin)), "l"(src), "r"(bytes), "r"(rsmem_u32(&bar)) : "memory");
    }
    if (do_small) {                                       // the small per-agent arrays ride along while the copy is in flight
        const int A = P.B.act_dim;
        for (int k = threadIdx.x; k < A * n_a; k += PUSH_THREADS) {
            const int a = k / A, d = k - a * A;
            const size_t s_ = (size_t)e * A * n_a + (size_t)d * n_a + a;
            P.B.act[row * A + k] = ldf(P.act, P.act_f32, s_);
            if (P.prior && P.B.act_prior) P.B.act_prior[row * A + k] = reinterpret_cast<const float *>(P.prior)[s_];
        }
        for (int a = threadIdx.x; a < n_a; a += PUSH_THREADS) {
            const size_t s_ = (size_t)e * n_a + a;
            P.B.rew[row + a] = reinterpret_cast<const float *>(P.rew)[s_];
            P.B.done[row + a] = P.done[s_] ? 1.0f : 0.0f;
            if (P.log_pi && P.B.log_pi) P.B.log_pi[row + a] = P.log_pi[s_];
        }
    }
    if (pass < 0) return;                                 // small arrays only
    __syncthreads();                                      // barrier initialised before anyone polls it
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "RWAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t"
        "@P1 bra RDONE;\n\t"
        "bra RWAIT;\n\t"
        "RDONE:\n\t}" ::"r"(rsmem_u32(&bar)) : "memory");
    {   // out[a][d] = in[d][a]; consecutive threads write consecutive words, reads stride n_a (2-way conflicts at n_a = 30)
        int a = threadIdx.x / D, d = threadIdx.x - a * D;
        const int sa = PUSH_THREADS / D, sd = PUSH_THREADS - sa * D;
#pragma unroll 4
        for (int k = threadIdx.x; k < n; k += PUSH_THREADS) {
            tout[k] = tin[d * n_a + a];
            a += sa; d += sd;
            if (d >= D) { d -= D; ++a; }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes visible to the bulk store
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(rsmem_u32(tout)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must stay alive until it has been read
    }
}

// The small per-agent arrays alone (time-indexed ring: the observations are written elsewhere): one thread per agent row.
__global__ void __launch_bounds__(256) k_rollout_push_small(const PushParams P, long n_rows) {
    const int span = P.a1 - P.a0, A = P.B.act_dim, n_a = P.n_a;
    for (long r = blockIdx.x * (long)blockDim.x + threadIdx.x; r < n_rows; r += (long)gridDim.x * blockDim.x) {
        const long e = r / span; const int a = P.a0 + (int)(r - e * span);
        const long row = P.row0 + r;
        for (int d = 0; d < A; ++d) {
            const size_t s_ = (size_t)e * A * n_a + (size_t)d * n_a + a;
            P.B.act[row * A + d] = ldf(P.act, P.act_f32, s_);
            if (P.prior && P.B.act_prior) P.B.act_prior[row * A + d] = ldf(P.prior, P.out_f32, s_);
        }
        const size_t s1 = (size_t)e * n_a + a;
        P.B.rew[row] = ldf(P.rew, P.out_f32, s1);
        P.B.done[row] = P.done[s1] ? 1.0f : 0.0f;
        if (P.log_pi && P.B.log_pi) P.B.log_pi[row] = P.log_pi[s1];
    }
}

struct GatherParams {
    RolloutBuf B;
    const long *idx;           // [n] ring rows (device)
    int n;
    float *obs, *act, *rew, *next_obs, *done, *prior, *log_pi;   // [n][dim] outputs; prior / log_pi may be NULL
    long next_off;             // >= 0: next_obs of row r is row r + next_off of the OBS array (time-indexed ring); < 0: the next_obs array
    long capacity;             // rows of the ring: out-of-range indices are clamped into it (and flagged), never dereferenced
    int *bad;                  // device flag set when an index was out of range (may be NULL)
};

__global__ void __launch_bounds__(256) k_rollout_gather(const GatherParams P) {
    const long gw = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= P.n) return;
    const int warp = (int)gw;
    long r = P.idx[warp];
    const long hi = P.capacity - 1 - (P.next_off > 0 ? P.next_off : 0);        // last row whose next_obs row is still inside the ring
    if (r < 0 || r > hi) { if (lane == 0 && P.bad) *P.bad = 1; r = r < 0 ? 0 : (hi < 0 ? 0 : hi); }
    const int D = P.B.obs_dim, A = P.B.act_dim;
    const float *nsrc = P.next_off >= 0 ? P.B.obs + (r + P.next_off) * D : P.B.next_obs + r * D;
    if ((D & 3) == 0) {                                    // rows are 16-byte aligned: vector copies
        const float4 *s0 = reinterpret_cast<const float4 *>(P.B.obs + r * D), *s1 = reinterpret_cast<const float4 *>(nsrc);
        float4 *d0 = reinterpret_cast<float4 *>(P.obs + (size_t)warp * D), *d1 = reinterpret_cast<float4 *>(P.next_obs + (size_t)warp * D);
        for (int k = lane; k < D / 4; k += 32) { d0[k] = s0[k]; d1[k] = s1[k]; }
    } else {
        for (int k = lane; k < D; k += 32) { P.obs[(size_t)warp * D + k] = P.B.obs[r * D + k]; P.next_obs[(size_t)warp * D + k] = nsrc[k]; }
    }
    for (int k = lane; k < A; k += 32) {
        P.act[(size_t)warp * A + k] = P.B.act[r * A + k];
        if (P.prior) P.prior[(size_t)warp * A + k] = P.B.act_prior[r * A + k];
    }
    if (lane == 0) {
        P.rew[warp] = P.B.rew[r]; P.done[warp] = P.B.done[r];
        if (P.log_pi) P.log_pi[warp] = P.B.log_pi[r];
    }
}

}  // namespace swarm
