// mnk_hostloop.cu -- the end-to-end path for callers that hold HOST buffers, K steps per call.
//
// mnk_step_host (mnk_env.cu) is one step per call: H2D copy, launch, D2H copy, stream synchronise -- a 10 us kernel
// wrapped in 20-35 us of host round trip (tools/e2e_probe.py).  mnk_step_host_loop runs a whole sequence of dense
// env steps (TorchVectorMnkEnv.step, src/env/torch_vector_mnk_env.py:55-84, K times) whose actions are already in
// pinned host memory -- a recorded game trace, an evaluation script, a host-side policy that works a slab ahead --
// as a three-stage pipeline over SLABS of `slab_steps` steps:
//
//     copy-in stream :  H2D actions of slab i+1           (one cudaMemcpyAsync per slab, queued a slab ahead of the host)
//     caller's stream:  the steps of slab i               (ONE launch: mnk_step_slab loops over a tile's steps)
//     copy-out stream:  D2H rewards + dones of slab i-1   (one cudaMemcpyAsync per slab; optionally every
//                                                          step's f32 observation + bool mask as well)
//
// Device-side action / result slabs are double-buffered; cross-stream ordering is by events; the host blocks once
// per slab (cudaEventSynchronize on the copy-out of the PREVIOUS slab), which is also the point where a consumer
// may read that slab's results.  Streams and events live in a caller-owned `mnk_host_pipe` handle so that a loop
// of short calls does not pay for their creation; the library itself keeps no global state.
#include "mnk_dispatch.cuh"

struct mnk_host_pipe {
    cudaStream_t in = nullptr, out = nullptr;
    cudaEvent_t in_done[MNK_HOST_LOOP_MAX_BUFFERS] = {}, compute_done[MNK_HOST_LOOP_MAX_BUFFERS] = {}, out_done[MNK_HOST_LOOP_MAX_BUFFERS] = {};
    int device = -1;
};

static void pipe_free(mnk_host_pipe* p) {
    if (p == nullptr) return;
    for (int i = 0; i < MNK_HOST_LOOP_MAX_BUFFERS; ++i) {
        if (p->in_done[i]) cudaEventDestroy(p->in_done[i]);
        if (p->compute_done[i]) cudaEventDestroy(p->compute_done[i]);
        if (p->out_done[i]) cudaEventDestroy(p->out_done[i]);
    }
    if (p->in) cudaStreamDestroy(p->in);
    if (p->out) cudaStreamDestroy(p->out);
    delete p;
}

extern "C" {

int mnk_host_pipe_create(void** pipe) {
    if (pipe == nullptr) return MNK_ERR_NULL;
    *pipe = nullptr;
    mnk_host_pipe* p = new mnk_host_pipe();
    cudaError_t e = cudaGetDevice(&p->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->out, cudaStreamNonBlocking);
    for (int i = 0; i < MNK_HOST_LOOP_MAX_BUFFERS && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&p->in_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->compute_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->out_done[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        pipe_free(p);
        return (int)e;
    }
    *pipe = p;
    return MNK_OK;
}

int mnk_host_pipe_destroy(void* pipe) {
    pipe_free(static_cast<mnk_host_pipe*>(pipe));
    return MNK_OK;
}

int mnk_step_host_loop(const mnk_state_t* st, const mnk_host_loop_t* job, void* pipe, uint32_t flags, void* stream) {
    if (int rc = mnk_check_state(st)) return rc;
    if (job == nullptr || job->host_actions == nullptr || job->host_rd == nullptr || job->dev_actions == nullptr ||
        job->dev_rd == nullptr)
        return MNK_ERR_NULL;
    const int64_t K = job->steps, S = job->slab_steps;
    const int NB = job->buffers == 0 ? 2 : job->buffers;           // slabs in flight
    if (K < 0 || S < 1 || job->ring < 0 || NB < 2 || NB > MNK_HOST_LOOP_MAX_BUFFERS) return MNK_ERR_ARG;
    const bool want_views = job->ring > 0;
    if (want_views && (job->obs_ring == nullptr || job->mask_ring == nullptr)) return MNK_ERR_NULL;
    const bool copy_views = job->host_obs != nullptr || job->host_mask != nullptr;
    // a ring slot is rewritten `ring` steps later; with the views copied out, up to NB slabs are in flight
    if (copy_views && (!want_views || job->ring < NB * S)) return MNK_ERR_ARG;
    if (K == 0 || st->num_envs == 0) return MNK_OK;

    mnk_host_pipe* p = static_cast<mnk_host_pipe*>(pipe);
    const bool own_pipe = (p == nullptr);
    if (own_pipe) {
        void* made = nullptr;
        if (int rc = mnk_host_pipe_create(&made)) return rc;
        p = static_cast<mnk_host_pipe*>(made);
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t n = (size_t)st->num_envs;
    const size_t cells = (size_t)st->m * st->n;
    const size_t abytes = n * ((flags & MNK_STEP_ACTIONS_I32) ? 4 : 8);
    const size_t rbytes = 5 * n;
    const uint32_t step_flags = (flags & (MNK_STEP_ACTIONS_I32 | MNK_STEP_AUTORESET)) | MNK_STEP_PDL;
    const int64_t slabs = (K + S - 1) / S;
    int rc = MNK_OK;
    cudaError_t e = cudaSuccess;
    auto fail = [&](cudaError_t err) { if (rc == MNK_OK && err != cudaSuccess) rc = (int)err; return err != cudaSuccess; };

    // copy-in of slab j: its device buffer's previous tenant (slab j-2) must have been consumed by its kernels -- a
    // stream-side wait, so the copy is queued a whole slab ahead of the host (it used to be queued only after the host had
    // waited for slab j-2's results, which serialised copy-in behind copy-out: 140 us per 8-step slab instead of ~80)
    auto copy_in = [&](int64_t j) -> bool {
        if (j >= slabs) return true;
        const int b = (int)(j % NB);
        const int64_t t0 = j * S, cnt = (K - t0 < S) ? (K - t0) : S;
        char* d_act = static_cast<char*>(job->dev_actions) + (size_t)b * S * abytes;
        if (j >= NB && fail(cudaStreamWaitEvent(p->in, p->compute_done[b], 0))) return false;
        if (fail(cudaMemcpyAsync(d_act, static_cast<const char*>(job->host_actions) + (size_t)t0 * abytes, (size_t)cnt * abytes,
                                 cudaMemcpyHostToDevice, p->in)))
            return false;
        return !fail(cudaEventRecord(p->in_done[b], p->in));
    };
    for (int j = 0; j < NB - 1; ++j) copy_in(j);
    for (int64_t i = 0; i < slabs && rc == MNK_OK; ++i) {
        const int b = (int)(i % NB);
        const int64_t t0 = i * S, cnt = (K - t0 < S) ? (K - t0) : S;
        char* d_act = static_cast<char*>(job->dev_actions) + (size_t)b * S * abytes;
        char* d_rd = static_cast<char*>(job->dev_rd) + (size_t)b * S * rbytes;
        // ---- compute: needs the actions; its result slab's previous tenant must have left for the host
        if (fail(cudaStreamWaitEvent(s, p->in_done[b], 0))) break;
        if (i >= NB && fail(cudaStreamWaitEvent(s, p->out_done[b], 0))) break;
        // the slab's steps: ONE launch per MNK_MAX_SLAB_STEPS steps (mnk_step_slab: a tile's steps only depend on the same
        // tile, so the kernel loops over them), not one launch per step -- the host side of a slab is then a launch, two
        // copies and five event operations whatever the slab length
        for (int64_t j0 = 0; j0 < cnt && rc == MNK_OK; j0 += MNK_MAX_SLAB_STEPS) {
            const int part = (int)((cnt - j0 < MNK_MAX_SLAB_STEPS) ? (cnt - j0) : MNK_MAX_SLAB_STEPS);
            float* obs[MNK_MAX_SLAB_STEPS];
            uint8_t* mask[MNK_MAX_SLAB_STEPS];
            for (int j = 0; j < part; ++j) {
                const int64_t t = t0 + j0 + j;
                obs[j] = want_views ? job->obs_ring[t % job->ring] : nullptr;
                mask[j] = want_views ? job->mask_ring[t % job->ring] : nullptr;
            }
            rc = mnk_step_slab(st, d_act + (size_t)j0 * abytes, (int64_t)abytes, d_rd + (size_t)j0 * rbytes, (int64_t)rbytes, part,
                               obs, mask, step_flags, s);
        }
        if (rc != MNK_OK) break;
        if (fail(cudaEventRecord(p->compute_done[b], s))) break;
        // ---- copy-in NB - 1 slabs ahead (that buffer was last read by slab i-1, whose completion event is already recorded)
        if (!copy_in(i + NB - 1)) break;
        // ---- copy-out
        if (fail(cudaStreamWaitEvent(p->out, p->compute_done[b], 0))) break;
        if (fail(cudaMemcpyAsync(static_cast<char*>(job->host_rd) + (size_t)t0 * rbytes, d_rd, (size_t)cnt * rbytes,
                                 cudaMemcpyDeviceToHost, p->out)))
            break;
        if (copy_views) {
            for (int64_t j = 0; j < cnt && rc == MNK_OK; ++j) {
                const int64_t t = t0 + j;
                if (job->host_obs != nullptr)
                    fail(cudaMemcpyAsync(job->host_obs + (size_t)t * n * 2 * cells, job->obs_ring[t % job->ring],
                                         n * 2 * cells * sizeof(float), cudaMemcpyDeviceToHost, p->out));
                if (job->host_mask != nullptr)
                    fail(cudaMemcpyAsync(job->host_mask + (size_t)t * n * cells, job->mask_ring[t % job->ring], n * cells,
                                         cudaMemcpyDeviceToHost, p->out));
            }
            if (rc != MNK_OK) break;
        }
        if (fail(cudaEventRecord(p->out_done[b], p->out))) break;
        // ---- the one host wait per slab: the results of slab i - (NB - 1) are now in host memory
        if (i >= NB - 1 && fail(cudaEventSynchronize(p->out_done[(i - (NB - 1)) % NB]))) break;
    }
    // drain: last slab's copy-out (also orders everything before the caller's next use of its stream / buffers)
    e = cudaStreamSynchronize(p->out);
    if (rc == MNK_OK && e != cudaSuccess) rc = (int)e;
    if (rc != MNK_OK) {   // leave nothing in flight that references the caller's buffers
        cudaStreamSynchronize(p->in);
        cudaStreamSynchronize(s);
    }
    if (own_pipe) pipe_free(p);
    return rc;
}

}  // extern "C"
