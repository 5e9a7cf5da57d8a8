// particlesolver_b200/csrc/ps_stream_io.cu — host-buffer steps with the PCIe transfers hidden behind the solver.
//
// ps_step_streamed is ParticleSystem::update for a host that owns positions and velocities (a host-side viewer or coupling code:
// the reference keeps its positions in a GL buffer and reads velocities back through copyArrayFromDevice, particlesystem.cpp:
// 122-142, 248-262): every call takes the step's inputs from host memory and delivers its result to host memory.  Done naively
// (upload, step, download on one stream) the two 16 B/particle transfers per direction serialise with the step.  Here the context
// keeps two staging frames per direction and two copy streams:
//     h2d stream      H2D host inputs of call k+1 -> staging_in[(k+1) & 1]            (while call k computes)
//     solver stream   staging_in[k & 1] -> pos, vel (D2D);  ps_step;  pos, vel -> staging_out[k & 1] (D2D)
//     d2h stream      staging_out[k & 1] -> host outputs of call k                    (while call k+1 computes)
// chained by events, so PCIe runs full duplex beside the kernels and a step costs max(solver, copy) instead of their sum.  The
// D2D copies are 64 B/particle of HBM traffic (~25 us at 1M particles).  Nothing here blocks the host; ps_io_wait does.
#include "ps_context.h"

#define IO_CU(call)                                                                                          \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) { ps_set_error("%s: %s", #call, cudaGetErrorString(e_)); return PS_ERR_CUDA; } \
    } while (0)

namespace {
struct DevGuard {
    int prev = 0;
    explicit DevGuard(int d) { cudaGetDevice(&prev); if (prev != d) cudaSetDevice(d); else prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

static int io_ensure(PsCtx *c, uint64_t n) {
    PsStreamIo &io = c->io;
    if (!io.h2d) {
        IO_CU(cudaStreamCreateWithFlags(&io.h2d, cudaStreamNonBlocking));
        IO_CU(cudaStreamCreateWithFlags(&io.d2h, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) {
            IO_CU(cudaEventCreateWithFlags(&io.in_ready[k], cudaEventDisableTiming));
            IO_CU(cudaEventCreateWithFlags(&io.in_free[k], cudaEventDisableTiming));
            IO_CU(cudaEventCreateWithFlags(&io.out_ready[k], cudaEventDisableTiming));
            IO_CU(cudaEventCreateWithFlags(&io.out_done[k], cudaEventDisableTiming));
        }
    }
    if (io.cap < n) {
        IO_CU(cudaStreamSynchronize(io.h2d));
        IO_CU(cudaStreamSynchronize(io.d2h));
        IO_CU(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < 2; k++)
            for (int a = 0; a < 2; a++) {
                if (io.in[k][a]) IO_CU(cudaFree(io.in[k][a]));
                if (io.out[k][a]) IO_CU(cudaFree(io.out[k][a]));
                IO_CU(cudaMalloc((void **)&io.in[k][a], n * sizeof(float4)));
                IO_CU(cudaMalloc((void **)&io.out[k][a], n * sizeof(float4)));
            }
        io.cap = n;
        io.calls = 0;
        io.begun = false;
        for (int k = 0; k < 2; k++) { io.prefetched[k] = false; io.frames_used[k] = false; }
    }
    return PS_OK;
}

void ps_io_free(PsCtx *c) {
    PsStreamIo &io = c->io;
    if (!io.h2d) return;
    cudaStreamSynchronize(io.h2d);
    cudaStreamSynchronize(io.d2h);
    for (int k = 0; k < 2; k++) {
        for (int a = 0; a < 2; a++) { if (io.in[k][a]) cudaFree(io.in[k][a]); if (io.out[k][a]) cudaFree(io.out[k][a]); }
        cudaEventDestroy(io.in_ready[k]); cudaEventDestroy(io.in_free[k]); cudaEventDestroy(io.out_ready[k]); cudaEventDestroy(io.out_done[k]);
    }
    cudaStreamDestroy(io.h2d);
    cudaStreamDestroy(io.d2h);
    io = PsStreamIo{};
}

// The two halves of a streamed step, for callers that issue the step themselves (the slab-decomposed step is a sequence of stage
// calls and exchanges, particlesolver_b200/slab.py): ps_io_begin stages this step's inputs (H2D on the copy stream, then D2D into
// pos / vel on the solver stream), ps_io_end delivers the result (D2D to the staging frame on the solver stream, D2H on the copy
// stream).  Both cover the OWNED particles (ghost copies of a slab context are rebuilt by the halo exchange).  One begin and one
// end per step, in that order; either may be skipped by passing NULL pointers.
extern "C" int ps_io_begin(PsCtx *c, const float *pos_in, const float *vel_in) {
    if (!c) { ps_set_error("null context"); return PS_ERR_INVALID; }
    const uint64_t n = c->n - c->n_ghost;
    if (!n || (!pos_in && !vel_in)) return PS_OK;
    DevGuard dg(c->device);
    int r = io_ensure(c, c->capacity ? c->capacity : n);
    if (r != PS_OK) return r;
    PsStreamIo &io = c->io;
    const int k = (int)(io.calls & 1);
    const size_t bytes = n * sizeof(float4);
    const float *src[2] = {pos_in, vel_in};
    float4 *state[2] = {c->pos, c->vel};
    // host -> staging frame k on the h2d stream, once the frame's previous contents have been consumed — unless ps_io_prefetch has
    // already put exactly these inputs there
    const bool staged = io.prefetched[k] && io.prefetch_n[k] == n && io.prefetch_src[k][0] == pos_in && io.prefetch_src[k][1] == vel_in;
    io.prefetched[k] = false;
    if (!staged) {
        if (io.frames_used[k]) IO_CU(cudaStreamWaitEvent(io.h2d, io.in_free[k], 0));
        for (int a = 0; a < 2; a++)
            if (src[a]) IO_CU(cudaMemcpyAsync(io.in[k][a], src[a], bytes, cudaMemcpyHostToDevice, io.h2d));
        IO_CU(cudaEventRecord(io.in_ready[k], io.h2d));
    }
    io.frames_used[k] = true;
    io.begun = true;
    IO_CU(cudaStreamWaitEvent(c->stream, io.in_ready[k], 0));
    for (int a = 0; a < 2; a++)
        if (src[a]) IO_CU(cudaMemcpyAsync(state[a], io.in[k][a], bytes, cudaMemcpyDeviceToDevice, c->stream));
    IO_CU(cudaEventRecord(io.in_free[k], c->stream));
    if (pos_in) c->grid_valid = false;
    return PS_OK;
}

// Issues the host -> device transfer of the NEXT step's inputs now (into the staging frame that step's ps_io_begin will commit), so
// that it overlaps the step in progress even when the step's issue blocks the host (a slab step waits for its neighbours' counts).
// The following ps_io_begin must be given the same pointers; if the owned particle count has changed in between it simply repeats
// the transfer.
extern "C" int ps_io_prefetch(PsCtx *c, const float *pos_in, const float *vel_in) {
    if (!c) { ps_set_error("null context"); return PS_ERR_INVALID; }
    const uint64_t n = c->n - c->n_ghost;
    if (!n || (!pos_in && !vel_in)) return PS_OK;
    DevGuard dg(c->device);
    int r = io_ensure(c, c->capacity ? c->capacity : n);
    if (r != PS_OK) return r;
    PsStreamIo &io = c->io;
    const int k = (int)((io.calls + (io.begun ? 1 : 0)) & 1);   // the frame of the next ps_io_begin
    const size_t bytes = n * sizeof(float4);
    const float *src[2] = {pos_in, vel_in};
    if (io.frames_used[k]) IO_CU(cudaStreamWaitEvent(io.h2d, io.in_free[k], 0));
    for (int a = 0; a < 2; a++)
        if (src[a]) IO_CU(cudaMemcpyAsync(io.in[k][a], src[a], bytes, cudaMemcpyHostToDevice, io.h2d));
    IO_CU(cudaEventRecord(io.in_ready[k], io.h2d));
    io.prefetched[k] = true; io.prefetch_n[k] = n; io.prefetch_src[k][0] = pos_in; io.prefetch_src[k][1] = vel_in;
    return PS_OK;
}

extern "C" int ps_io_end(PsCtx *c, float *pos_out, float *vel_out) {
    if (!c) { ps_set_error("null context"); return PS_ERR_INVALID; }
    const uint64_t n = c->n - c->n_ghost;
    DevGuard dg(c->device);
    int r = io_ensure(c, c->capacity ? c->capacity : (n ? n : 1));
    if (r != PS_OK) return r;
    PsStreamIo &io = c->io;
    const int k = (int)(io.calls & 1);
    const size_t bytes = n * sizeof(float4);
    float *dst[2] = {pos_out, vel_out};
    float4 *state[2] = {c->pos, c->vel};
    // state -> staging frame k on the solver stream (after the frame's previous download), then to the host on the d2h stream
    if (n && (pos_out || vel_out)) {
        if (io.calls >= 2) IO_CU(cudaStreamWaitEvent(c->stream, io.out_done[k], 0));
        for (int a = 0; a < 2; a++)
            if (dst[a]) IO_CU(cudaMemcpyAsync(io.out[k][a], state[a], bytes, cudaMemcpyDeviceToDevice, c->stream));
        IO_CU(cudaEventRecord(io.out_ready[k], c->stream));
        IO_CU(cudaStreamWaitEvent(io.d2h, io.out_ready[k], 0));
        for (int a = 0; a < 2; a++)
            if (dst[a]) IO_CU(cudaMemcpyAsync(dst[a], io.out[k][a], bytes, cudaMemcpyDeviceToHost, io.d2h));
    }
    IO_CU(cudaEventRecord(io.out_done[k], io.d2h));
    io.calls++;
    io.begun = false;
    return PS_OK;
}

extern "C" int ps_step_streamed(PsCtx *c, float dt, const float *pos_in, const float *vel_in, float *pos_out, float *vel_out) {
    if (!c) { ps_set_error("null context"); return PS_ERR_INVALID; }
    if (c->n_ghost) { ps_set_error("ps_step_streamed: a slab context steps through its stage calls; use ps_io_begin / ps_io_end around them"); return PS_ERR_STATE; }
    int r = ps_io_begin(c, pos_in, vel_in);
    if (r != PS_OK) return r;
    if ((r = ps_step(c, dt)) != PS_OK) return r;
    return ps_io_end(c, pos_out, vel_out);
}

// Blocks until the outputs of the ps_step_streamed call made `calls_back` calls ago (0 = the last one, 1 = the one before) are in
// host memory.  Only the last two calls can still be in flight.
extern "C" int ps_io_wait(PsCtx *c, uint32_t calls_back) {
    if (!c) { ps_set_error("null context"); return PS_ERR_INVALID; }
    PsStreamIo &io = c->io;
    if (!io.h2d || io.calls <= calls_back) return PS_OK;
    DevGuard dg(c->device);
    if (calls_back >= 2) return PS_OK;  // older calls completed before their frame was reused
    const int k = (int)((io.calls - 1 - calls_back) & 1);
    IO_CU(cudaEventSynchronize(io.out_done[k]));
    return PS_OK;
}
