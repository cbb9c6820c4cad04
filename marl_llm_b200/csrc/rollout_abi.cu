// rollout_abi.cu — C ABI of the device-resident replay storage (include/swarm_b200.h, group 3).  No CPU implementation.
#include "rollout_kernels.cuh"
#include "../../include/swarm_b200.h"

#include <algorithm>
#include <string>

using namespace swarm;

extern "C" int swarm_set_last_error_(int code, const char *msg);   // swarm_abi.cu

namespace {
int rfail(int code, const std::string &m) { return swarm_set_last_error_(code, m.c_str()); }
#define RCU_TRY(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t err__ = (expr);                                                                     \
        if (err__ != cudaSuccess) return rfail(SWARM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__)); \
    } while (0)

int check_buf(const swarm_rollout_buffers *b) {
    if (!b) return rfail(SWARM_ERR_INVALID, "null rollout buffers");
    if (b->struct_size != (int32_t)sizeof(swarm_rollout_buffers)) return rfail(SWARM_ERR_INVALID, "swarm_rollout_buffers struct_size mismatch (ABI)");
    if (b->obs_dim <= 0 || b->act_dim <= 0 || b->capacity <= 0) return rfail(SWARM_ERR_INVALID, "rollout dims must be positive");
    if (!b->obs || !b->act || !b->rew || !b->done) return rfail(SWARM_ERR_INVALID, "required rollout buffer is NULL");
    return SWARM_OK;
}
RolloutBuf to_dev(const swarm_rollout_buffers *b) {
    RolloutBuf B;
    B.obs = b->obs; B.act = b->act; B.act_prior = b->act_prior; B.log_pi = b->log_pi; B.rew = b->rew; B.next_obs = b->next_obs; B.done = b->done;
    B.obs_dim = b->obs_dim; B.act_dim = b->act_dim;
    return B;
}
}  // namespace

extern "C" {

int swarm_rollout_push(const swarm_rollout_buffers *buf, int64_t row0, int32_t num_envs, int32_t n_a, int32_t agent_start,
                       int32_t agent_stop, const void *obs, const void *next_obs, const void *reward, const uint8_t *done,
                       const void *act_prior, int out_dtype, const void *act, int act_dtype, const float *log_pi, void *stream) {
    return swarm_rollout_push_parts(buf, row0, num_envs, n_a, agent_start, agent_stop, obs, next_obs, reward, done, act_prior, out_dtype,
                                    act, act_dtype, log_pi, SWARM_PUSH_OBS | SWARM_PUSH_NEXT_OBS | SWARM_PUSH_SMALL, stream);
}

int swarm_rollout_push_parts(const swarm_rollout_buffers *buf, int64_t row0, int32_t num_envs, int32_t n_a, int32_t agent_start,
                             int32_t agent_stop, const void *obs, const void *next_obs, const void *reward, const uint8_t *done,
                             const void *act_prior, int out_dtype, const void *act, int act_dtype, const float *log_pi, int parts,
                             void *stream) {
    int rc = check_buf(buf);
    if (rc != SWARM_OK) return rc;
    if (parts <= 0 || parts > 7) return rfail(SWARM_ERR_INVALID, "parts must be a non-empty combination of SWARM_PUSH_*");
    if ((parts & SWARM_PUSH_NEXT_OBS) && !buf->next_obs) return rfail(SWARM_ERR_INVALID, "this ring has no next_obs array (time-indexed ring)");
    if (((parts & SWARM_PUSH_OBS) && !obs) || ((parts & SWARM_PUSH_NEXT_OBS) && !next_obs) ||
        ((parts & SWARM_PUSH_SMALL) && (!reward || !done || !act)))
        return rfail(SWARM_ERR_INVALID, "null push input");
    if (num_envs <= 0 || n_a <= 0 || agent_start < 0 || agent_stop > n_a || agent_stop <= agent_start)
        return rfail(SWARM_ERR_INVALID, "bad env count / agent slice");
    if ((out_dtype != SWARM_F32 && out_dtype != SWARM_F64) || (act_dtype != SWARM_F32 && act_dtype != SWARM_F64))
        return rfail(SWARM_ERR_INVALID, "bad dtype");
    const int64_t rows = (int64_t)num_envs * (agent_stop - agent_start);
    if (row0 < 0 || row0 + rows > buf->capacity) return rfail(SWARM_ERR_INVALID, "push does not fit the ring at row0");
    if (act_prior && !buf->act_prior) return rfail(SWARM_ERR_INVALID, "act_prior given but the ring has no act_prior array");
    PushParams P;
    P.B = to_dev(buf); P.row0 = (long)row0; P.n_a = n_a; P.a0 = agent_start; P.a1 = agent_stop;
    P.obs = obs; P.next_obs = next_obs; P.rew = reward; P.prior = act_prior; P.act = act; P.log_pi = log_pi; P.done = done;
    P.out_f32 = (out_dtype == SWARM_F32); P.act_f32 = (act_dtype == SWARM_F32); P.parts = parts;
    if (parts == SWARM_PUSH_SMALL) {
        const int blocks = (int)std::min<int64_t>((rows + 255) / 256, 148 * 16);
        k_rollout_push_small<<<blocks, 256, 0, (cudaStream_t)stream>>>(P, (long)rows);
        RCU_TRY(cudaGetLastError());
        return SWARM_OK;
    }
    const size_t tile_bytes = (size_t)buf->obs_dim * n_a * sizeof(float);
    const bool tma = P.out_f32 && agent_start == 0 && agent_stop == n_a && (tile_bytes % 16) == 0 && 2 * tile_bytes <= 100 * 1024 &&
                     (!(parts & 1) || ((uintptr_t)obs % 16) == 0) && (!(parts & 2) || ((uintptr_t)next_obs % 16) == 0) && ((uintptr_t)buf->obs % 16) == 0 &&
                     (!(parts & 2) || ((uintptr_t)buf->next_obs % 16) == 0) && ((size_t)row0 * buf->obs_dim * sizeof(float)) % 16 == 0 &&
                     ((size_t)n_a * buf->obs_dim * sizeof(float)) % 16 == 0;
    if (tma) {      // whole-env contiguous fp32 tiles: bulk copy in, transpose in shared memory, bulk copy out
        const size_t sm = 2 * tile_bytes;
        RCU_TRY(cudaFuncSetAttribute((const void *)k_rollout_push_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        const unsigned npass = (unsigned)((parts & 1) + ((parts >> 1) & 1));
        k_rollout_push_tma<<<dim3((unsigned)num_envs, npass ? npass : 1), PUSH_THREADS, sm, (cudaStream_t)stream>>>(P);
        RCU_TRY(cudaGetLastError());
        return SWARM_OK;
    }
    const size_t smem = (size_t)buf->obs_dim * (PUSH_CHUNK + 1) * sizeof(float);
    if (smem > 48 * 1024) RCU_TRY(cudaFuncSetAttribute((const void *)k_rollout_push, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid((unsigned)num_envs, (unsigned)((agent_stop - agent_start + PUSH_CHUNK - 1) / PUSH_CHUNK));
    k_rollout_push<<<grid, PUSH_THREADS, smem, (cudaStream_t)stream>>>(P);
    RCU_TRY(cudaGetLastError());
    return SWARM_OK;
}

int swarm_rollout_gather(const swarm_rollout_buffers *buf, const int64_t *rows_dev, int32_t n, float *obs, float *act, float *reward,
                         float *next_obs, float *done, float *act_prior, float *log_pi, void *stream) {
    return swarm_rollout_gather_ring(buf, rows_dev, n, -1, obs, act, reward, next_obs, done, act_prior, log_pi, stream);
}

int swarm_rollout_gather_ring(const swarm_rollout_buffers *buf, const int64_t *rows_dev, int32_t n, int64_t next_row_offset, float *obs,
                              float *act, float *reward, float *next_obs, float *done, float *act_prior, float *log_pi, void *stream) {
    int rc = check_buf(buf);
    if (rc != SWARM_OK) return rc;
    if (!rows_dev || n <= 0 || !obs || !act || !reward || !next_obs || !done) return rfail(SWARM_ERR_INVALID, "bad gather argument");
    if (next_row_offset < 0 && !buf->next_obs) return rfail(SWARM_ERR_INVALID, "this ring has no next_obs array: pass next_row_offset");
    if ((act_prior && !buf->act_prior) || (log_pi && !buf->log_pi)) return rfail(SWARM_ERR_INVALID, "optional array requested but not stored");
    GatherParams G;
    G.B = to_dev(buf); G.idx = reinterpret_cast<const long *>(rows_dev); G.n = n;
    G.obs = obs; G.act = act; G.rew = reward; G.next_obs = next_obs; G.done = done; G.prior = act_prior; G.log_pi = log_pi;
    G.next_off = (long)next_row_offset;
    G.capacity = (long)buf->capacity; G.bad = nullptr;      // rows outside [0, capacity) are clamped by the kernel, never read out of bounds
    k_rollout_gather<<<(unsigned)(((long)n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(G);
    RCU_TRY(cudaGetLastError());
    return SWARM_OK;
}

}  // extern "C"
