// srt_dispatch.cu — NCCL stream dispatcher behind include/srt_dispatch.h (libspleeterrt_dispatch.so).
// The schedule (who sends which stream to whom, in which group) is a pure host function shared by the NCCL executor and the
// CPU unit test; the executor only walks it.
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/srt_dispatch.h"

namespace {

thread_local std::string g_derr;
int dfail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_derr = buf;
    return code;
}
#define DCK(call)                                                                                                   \
    do {                                                                                                            \
        cudaError_t e_ = (call);                                                                                    \
        if (e_ != cudaSuccess) return dfail(SRT_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define NCK(call)                                                                                                   \
    do {                                                                                                            \
        ncclResult_t r_ = (call);                                                                                   \
        if (r_ != ncclSuccess) return dfail(SRT_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(r_)); \
    } while (0)

struct Guard {   // leave the caller's current device as it was
    int prev = -1;
    explicit Guard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct Op {
    int group, kind, peer, stream, slot, count;
};

std::vector<int> local_streams(int world, int rank, int n_streams)
{
    std::vector<int> ids;
    for (int i = rank; i < n_streams; i += world) ids.push_back(i);
    return ids;
}
void chunk_range(int n_local, int chunks, int c, int& a, int& b)
{
    const int per = (n_local + chunks - 1) / chunks;
    a = c * per < n_local ? c * per : n_local;
    b = (c + 1) * per < n_local ? (c + 1) * per : n_local;
}

// Issue order on every rank: scatter groups 0..chunks-1, then gather groups chunks..2*chunks-1.  Within a group root walks its
// peers in rank order and each peer's chunk in local order; the peer walks the same chunk in the same order: per (sender,
// receiver) pair the two lists agree element by element.
std::vector<Op> build_schedule(int world, int rank, int root, const size_t* n_samples, int n_streams, int pairs, int chunks)
{
    std::vector<Op> ops;
    for (int phase = 0; phase < 2; phase++)
        for (int c = 0; c < chunks; c++) {
            const int group = phase * chunks + c;
            for (int p = 0; p < world; p++) {
                if (p == root) continue;
                if (rank != root && rank != p) continue;
                const std::vector<int> ids = local_streams(world, p, n_streams);
                int a, b;
                chunk_range((int)ids.size(), chunks, c, a, b);
                for (int k = a; k < b; k++) {
                    const int i = ids[k], n = (int)n_samples[i];
                    if (phase == 0) {
                        for (int slot = 0; slot < 2; slot++) ops.push_back(Op{group, rank == root ? 0 : 1, rank == root ? p : root, i, slot, n});
                    } else {
                        for (int slot = 2; slot < 2 + 2 * pairs; slot++) ops.push_back(Op{group, rank == root ? 1 : 0, rank == root ? p : root, i, slot, n});
                    }
                }
            }
        }
    return ops;
}

}  // namespace

struct srt_dispatch {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0, device = 0;
    cudaStream_t comm_stream = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_out;
    cudaEvent_t ev_start = nullptr;
    float *d_in = nullptr, *d_out = nullptr;      // staging of a non-root rank: its streams' PCM, its results
    size_t in_cap = 0, out_cap = 0;
    float* d_bcast = nullptr;
    size_t bcast_cap = 0;
    // peer-memory mode: root's buffers as this rank sees them (own allocation on root, IPC mappings elsewhere)
    float *peer_in = nullptr, *peer_out = nullptr;
    size_t peer_in_floats = 0, peer_out_floats = 0;
    int peer_root = -1;
    bool peer_mapped = false;      // the pointers are IPC mappings (close, not free)
    float* d_flag = nullptr;       // the one-word all-reduce buffer
    cudaEvent_t ev_done = nullptr;
};

static void release_peer(srt_dispatch* d)
{
    if (d->peer_in) { if (d->peer_mapped) cudaIpcCloseMemHandle(d->peer_in); else cudaFree(d->peer_in); }
    if (d->peer_out) { if (d->peer_mapped) cudaIpcCloseMemHandle(d->peer_out); else cudaFree(d->peer_out); }
    d->peer_in = d->peer_out = nullptr;
    d->peer_in_floats = d->peer_out_floats = 0;
    d->peer_root = -1;
}

extern "C" const char* srt_dispatch_last_error(void) { return g_derr.c_str(); }

extern "C" int srt_dispatch_get_id(unsigned char id[SRT_DISPATCH_ID_BYTES])
{
    static_assert(sizeof(ncclUniqueId) == SRT_DISPATCH_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    NCK(ncclGetUniqueId(&u));
    std::memcpy(id, &u, sizeof u);
    return 0;
}

extern "C" int srt_dispatch_create(const unsigned char id[SRT_DISPATCH_ID_BYTES], int world, int rank, int device, srt_dispatch** out)
{
    if (!id || !out || world < 1 || rank < 0 || rank >= world) return dfail(SRT_ERR_ARG, "bad argument");
    *out = nullptr;
    Guard g(device);
    srt_dispatch* d = new srt_dispatch();
    d->world = world; d->rank = rank; d->device = device;
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof u);
    ncclResult_t r = ncclCommInitRank(&d->comm, world, u, rank);
    if (r != ncclSuccess) { delete d; return dfail(SRT_ERR_CUDA, "ncclCommInitRank: %s", ncclGetErrorString(r)); }
    if (cudaStreamCreateWithFlags(&d->comm_stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&d->ev_start, cudaEventDisableTiming) != cudaSuccess) {
        srt_dispatch_destroy(d);
        return dfail(SRT_ERR_CUDA, "stream / event creation failed");
    }
    *out = d;
    return 0;
}

extern "C" void srt_dispatch_destroy(srt_dispatch* d)
{
    if (!d) return;
    Guard g(d->device);
    if (d->comm_stream) cudaStreamSynchronize(d->comm_stream);
    for (cudaEvent_t e : d->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : d->ev_out) cudaEventDestroy(e);
    if (d->ev_start) cudaEventDestroy(d->ev_start);
    if (d->d_in) cudaFree(d->d_in);
    if (d->d_out) cudaFree(d->d_out);
    if (d->d_bcast) cudaFree(d->d_bcast);
    release_peer(d);
    if (d->d_flag) cudaFree(d->d_flag);
    if (d->ev_done) cudaEventDestroy(d->ev_done);
    if (d->comm) ncclCommDestroy(d->comm);
    if (d->comm_stream) cudaStreamDestroy(d->comm_stream);
    delete d;
}

static int bcast_bytes(srt_dispatch* d, void* host, size_t bytes, int root)
{
    Guard g(d->device);
    if (bytes > d->bcast_cap) {
        if (d->d_bcast) cudaFree(d->d_bcast);
        d->d_bcast = nullptr; d->bcast_cap = 0;
        DCK(cudaMalloc((void**)&d->d_bcast, bytes));
        d->bcast_cap = bytes;
    }
    if (d->rank == root) DCK(cudaMemcpyAsync(d->d_bcast, host, bytes, cudaMemcpyHostToDevice, d->comm_stream));
    NCK(ncclBroadcast(d->d_bcast, d->d_bcast, bytes, ncclChar, root, d->comm, d->comm_stream));
    if (d->rank != root) DCK(cudaMemcpyAsync(host, d->d_bcast, bytes, cudaMemcpyDeviceToHost, d->comm_stream));
    DCK(cudaStreamSynchronize(d->comm_stream));
    return 0;
}

extern "C" int srt_dispatch_broadcast_weights(srt_dispatch* d, float* coeffs, int n_nets, int root)
{
    if (!d || !coeffs || n_nets < 1 || root < 0 || root >= d->world) return dfail(SRT_ERR_ARG, "bad argument");
    int r = bcast_bytes(d, coeffs, (size_t)n_nets * SRT_COEFF_FLOATS * sizeof(float), root);
    if (d->d_bcast) { Guard g(d->device); cudaFree(d->d_bcast); d->d_bcast = nullptr; d->bcast_cap = 0; }   // 40 MB per net: not kept
    return r;
}

extern "C" int srt_dispatch_broadcast_sizes(srt_dispatch* d, size_t* values, int n, int root)
{
    if (!d || !values || n < 1 || root < 0 || root >= d->world) return dfail(SRT_ERR_ARG, "bad argument");
    return bcast_bytes(d, values, (size_t)n * sizeof(size_t), root);
}

extern "C" int srt_dispatch_local_streams(int world, int rank, int n_streams, int* ids)
{
    if (world < 1 || rank < 0 || rank >= world || n_streams < 0) return SRT_ERR_ARG;
    const std::vector<int> v = local_streams(world, rank, n_streams);
    if (ids) std::memcpy(ids, v.data(), v.size() * sizeof(int));
    return (int)v.size();
}

extern "C" long long srt_dispatch_schedule(int world, int rank, int root, const size_t* n_samples, int n_streams, int pairs, int chunks,
                                           int* rows, long long cap_rows)
{
    if (world < 1 || rank < 0 || rank >= world || root < 0 || root >= world || !n_samples || n_streams < 1 || pairs < 1 || chunks < 1)
        return dfail(SRT_ERR_ARG, "bad argument");
    const std::vector<Op> ops = build_schedule(world, rank, root, n_samples, n_streams, pairs, chunks);
    if (rows && cap_rows >= (long long)ops.size())
        for (size_t k = 0; k < ops.size(); k++) {
            const Op& o = ops[k];
            const int row[6] = {o.group, o.kind, o.peer, o.stream, o.slot, o.count};
            std::memcpy(rows + 6 * k, row, sizeof row);
        }
    return (long long)ops.size();
}

extern "C" void* srt_dispatch_comm_stream(srt_dispatch* d) { return d ? (void*)d->comm_stream : nullptr; }

extern "C" int srt_dispatch_wait(srt_dispatch* d)
{
    if (!d) return dfail(SRT_ERR_STATE, "null dispatcher");
    Guard g(d->device);
    DCK(cudaStreamSynchronize(d->comm_stream));
    return 0;
}

extern "C" int srt_dispatch_separate_device(srt_dispatch* d, srt_ctx* ctx, int root, const float* const* d_pcmL, const float* const* d_pcmR,
                                            const size_t* n_samples, int n_streams, const float* unaffected, float* const* d_out, int chunks)
{
    if (!d || !ctx || !n_samples || n_streams < 1 || root < 0 || root >= d->world) return dfail(SRT_ERR_ARG, "bad argument");
    if (chunks < 1) chunks = 1;
    const int pairs = srt_output_pairs(ctx);
    const bool is_root = d->rank == root;
    if (is_root && (!d_pcmL || !d_pcmR || !d_out)) return dfail(SRT_ERR_ARG, "root needs the stream and output pointers");
    Guard g(d->device);
    cudaStream_t cs = (cudaStream_t)srt_cuda_stream(ctx);
    const std::vector<int> mine = local_streams(d->world, d->rank, n_streams);
    const int n_local = (int)mine.size();
    // ---- staging of a non-root rank: [stream][L | R] in, [stream][pair][channel] out ---------------------------------------
    std::vector<size_t> off(n_local + 1, 0);
    for (int k = 0; k < n_local; k++) off[k + 1] = off[k] + ((n_samples[mine[k]] + 3) & ~(size_t)3);
    if (!is_root) {
        const size_t need_in = off[n_local] * 2, need_out = off[n_local] * 2 * pairs;
        if (need_in > d->in_cap) {
            DCK(cudaStreamSynchronize(d->comm_stream));
            if (d->d_in) cudaFree(d->d_in);
            d->d_in = nullptr; d->in_cap = 0;
            DCK(cudaMalloc((void**)&d->d_in, need_in * sizeof(float)));
            d->in_cap = need_in;
        }
        if (need_out > d->out_cap) {
            DCK(cudaStreamSynchronize(d->comm_stream));
            DCK(cudaStreamSynchronize(cs));
            if (d->d_out) cudaFree(d->d_out);
            d->d_out = nullptr; d->out_cap = 0;
            DCK(cudaMalloc((void**)&d->d_out, need_out * sizeof(float)));
            d->out_cap = need_out;
        }
    }
    while ((int)d->ev_in.size() < chunks) {
        cudaEvent_t a, b;
        DCK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        DCK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        d->ev_in.push_back(a);
        d->ev_out.push_back(b);
    }
    // pointers of local stream k as this rank's context sees them
    auto in_ptr = [&](int k, int ch) -> const float* {
        if (is_root) return ch ? d_pcmR[mine[k]] : d_pcmL[mine[k]];
        return d->d_in + off[k] * 2 + (ch ? ((n_samples[mine[k]] + 3) & ~(size_t)3) : 0);
    };
    auto out_ptr = [&](int k, int slot /* q * 2 + c */) -> float* {
        if (is_root) return d_out[(size_t)mine[k] * pairs * 2 + slot];
        return d->d_out + off[k] * 2 * pairs + (size_t)slot * ((n_samples[mine[k]] + 3) & ~(size_t)3);
    };
    // the transfers start after everything already queued on the compute stream (the caller's inputs may be produced there)
    DCK(cudaEventRecord(d->ev_start, cs));
    DCK(cudaStreamWaitEvent(d->comm_stream, d->ev_start, 0));
    const std::vector<Op> ops = build_schedule(d->world, d->rank, root, n_samples, n_streams, pairs, chunks);
    std::vector<int> local_index(n_streams, -1);
    for (int k = 0; k < n_local; k++) local_index[mine[k]] = k;
    size_t cur = 0;
    auto run_group = [&](int group) -> int {
        const size_t first = cur;
        while (cur < ops.size() && ops[cur].group == group) cur++;
        // NCCL fuses the point-to-point operations of a group; keep groups to a few hundred operations
        for (size_t a = first; a < cur; a += 512) {
            const size_t b = a + 512 < cur ? a + 512 : cur;
            NCK(ncclGroupStart());
            for (size_t k = a; k < b; k++) {
                const Op& o = ops[k];
                void* p;
                if (is_root) p = o.slot < 2 ? (void*)(o.slot ? d_pcmR[o.stream] : d_pcmL[o.stream]) : (void*)d_out[(size_t)o.stream * pairs * 2 + (o.slot - 2)];
                else p = o.slot < 2 ? (void*)in_ptr(local_index[o.stream], o.slot) : (void*)out_ptr(local_index[o.stream], o.slot - 2);
                if (o.kind == 0) NCK(ncclSend(p, (size_t)o.count, ncclFloat, o.peer, d->comm, d->comm_stream));
                else NCK(ncclRecv(p, (size_t)o.count, ncclFloat, o.peer, d->comm, d->comm_stream));
            }
            NCK(ncclGroupEnd());
        }
        return 0;
    };
    int r;
    // ---- scatter, chunk by chunk; the compute of chunk c waits only for its own inputs ------------------------------------------
    for (int c = 0; c < chunks; c++) {
        if ((r = run_group(c))) return r;
        if (!is_root) DCK(cudaEventRecord(d->ev_in[c], d->comm_stream));
    }
    // ---- separate the local share, chunk by chunk, on the context's stream ---------------------------------------------------------
    for (int c = 0; c < chunks; c++) {
        int a, b;
        chunk_range(n_local, chunks, c, a, b);
        if (a < b) {
            if (!is_root) DCK(cudaStreamWaitEvent(cs, d->ev_in[c], 0));
            std::vector<const float*> L(b - a), R(b - a);
            std::vector<size_t> n(b - a);
            std::vector<float*> o((size_t)(b - a) * pairs * 2);
            for (int k = a; k < b; k++) {
                L[k - a] = in_ptr(k, 0); R[k - a] = in_ptr(k, 1); n[k - a] = n_samples[mine[k]];
                for (int s = 0; s < pairs * 2; s++) o[(size_t)(k - a) * pairs * 2 + s] = out_ptr(k, s);
            }
            if (srt_separate_device(ctx, L.data(), R.data(), n.data(), b - a, unaffected, o.data()))
                return dfail(SRT_ERR_STATE, "rank %d chunk %d: %s", d->rank, c, srt_last_error());
        }
        DCK(cudaEventRecord(d->ev_out[c], cs));
    }
    // ---- gather, chunk by chunk: a chunk leaves as soon as it is computed --------------------------------------------------------
    for (int c = 0; c < chunks; c++) {
        if (!is_root) DCK(cudaStreamWaitEvent(d->comm_stream, d->ev_out[c], 0));
        if ((r = run_group(chunks + c))) return r;
    }
    if (is_root) {   // root's own share finishes on the compute stream: the comm stream (what srt_dispatch_wait waits on) joins it
        DCK(cudaStreamWaitEvent(d->comm_stream, d->ev_out[chunks - 1], 0));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------------
// peer-memory mode
// ---------------------------------------------------------------------------------------------------------------------------
extern "C" int srt_dispatch_peer_layout(const size_t* n_samples, int n_streams, int pairs, size_t* in_off, size_t* out_off,
                                        size_t* in_floats, size_t* out_floats)
{
    if (!n_samples || n_streams < 1 || pairs < 1) return dfail(SRT_ERR_ARG, "bad argument");
    size_t a = 0, b = 0;
    for (int i = 0; i < n_streams; i++) {
        const size_t np = (n_samples[i] + 3) & ~(size_t)3;
        if (in_off) in_off[i] = a;
        if (out_off) out_off[i] = b;
        a += 2 * np;
        b += (size_t)2 * pairs * np;
    }
    if (in_floats) *in_floats = a;
    if (out_floats) *out_floats = b;
    return 0;
}

extern "C" int srt_dispatch_peer_buffers(srt_dispatch* d, int root, size_t in_floats, size_t out_floats, float** d_in, float** d_out)
{
    if (!d || !d_in || !d_out || root < 0 || root >= d->world) return dfail(SRT_ERR_ARG, "bad argument");
    Guard g(d->device);
    DCK(cudaStreamSynchronize(d->comm_stream));
    release_peer(d);
    // {in_floats, out_floats, in handle, out handle} from root to everybody
    struct Msg {
        size_t in_floats, out_floats;
        cudaIpcMemHandle_t hin, hout;
    } msg;
    std::memset(&msg, 0, sizeof msg);
    if (d->rank == root) {
        if (in_floats < 1 || out_floats < 1) return dfail(SRT_ERR_ARG, "empty peer buffers");
        DCK(cudaMalloc((void**)&d->peer_in, in_floats * sizeof(float)));     // plain cudaMalloc: whole allocations, exportable
        DCK(cudaMalloc((void**)&d->peer_out, out_floats * sizeof(float)));
        d->peer_mapped = false;
        msg.in_floats = in_floats; msg.out_floats = out_floats;
        DCK(cudaIpcGetMemHandle(&msg.hin, d->peer_in));
        DCK(cudaIpcGetMemHandle(&msg.hout, d->peer_out));
    }
    int r = bcast_bytes(d, &msg, sizeof msg, root);
    if (r) return r;
    if (d->rank != root) {
        DCK(cudaIpcOpenMemHandle((void**)&d->peer_in, msg.hin, cudaIpcMemLazyEnablePeerAccess));
        DCK(cudaIpcOpenMemHandle((void**)&d->peer_out, msg.hout, cudaIpcMemLazyEnablePeerAccess));
        d->peer_mapped = true;
    }
    d->peer_in_floats = msg.in_floats; d->peer_out_floats = msg.out_floats; d->peer_root = root;
    if (!d->d_flag) DCK(cudaMalloc((void**)&d->d_flag, 2 * sizeof(float)));
    if (!d->ev_done) DCK(cudaEventCreateWithFlags(&d->ev_done, cudaEventDisableTiming));
    *d_in = d->peer_in;
    *d_out = d->peer_out;
    return 0;
}

extern "C" int srt_dispatch_separate_peer(srt_dispatch* d, srt_ctx* ctx, int root, const size_t* n_samples, int n_streams,
                                          const float* unaffected, int chunks)
{
    if (!d || !ctx || !n_samples || n_streams < 1 || root != d->peer_root) return dfail(SRT_ERR_ARG, "bad argument (srt_dispatch_peer_buffers first)");
    if (chunks < 1) chunks = 1;
    Guard g(d->device);
    const int pairs = srt_output_pairs(ctx);
    std::vector<size_t> in_off(n_streams), out_off(n_streams);
    size_t need_in = 0, need_out = 0;
    srt_dispatch_peer_layout(n_samples, n_streams, pairs, in_off.data(), out_off.data(), &need_in, &need_out);
    if (need_in > d->peer_in_floats || need_out > d->peer_out_floats) return dfail(SRT_ERR_CAPACITY, "batch exceeds the peer buffers");
    cudaStream_t cs = (cudaStream_t)srt_cuda_stream(ctx);
    const std::vector<int> mine = local_streams(d->world, d->rank, n_streams);
    const int n_local = (int)mine.size();
    // start line: a one-word all-reduce behind whatever every rank's compute stream already holds - on root that is the work
    // that produced the PCM (and consumed the previous batch's stems) - and in front of this batch's kernels
    DCK(cudaEventRecord(d->ev_start, cs));
    DCK(cudaStreamWaitEvent(d->comm_stream, d->ev_start, 0));
    NCK(ncclAllReduce(d->d_flag, d->d_flag + 1, 1, ncclFloat, ncclSum, d->comm, d->comm_stream));
    DCK(cudaEventRecord(d->ev_done, d->comm_stream));
    DCK(cudaStreamWaitEvent(cs, d->ev_done, 0));
    // chunks == 1 (or root): the overlap-add kernel stores straight into root's memory.  chunks > 1 on the other ranks: a chunk's
    // stems land in local memory first and leave with ONE copy-engine transfer per stream while the next chunk computes - a burst
    // of NVLink stores from seven GPUs at the end of everybody's step (they run in lockstep) is what limits the direct form.
    const bool staged = chunks > 1 && d->rank != root;
    std::vector<size_t> loff(n_local + 1, 0);
    if (staged) {
        for (int k = 0; k < n_local; k++) loff[k + 1] = loff[k] + (size_t)2 * pairs * ((n_samples[mine[k]] + 3) & ~(size_t)3);
        if (loff[n_local] > d->out_cap) {
            DCK(cudaStreamSynchronize(d->comm_stream));
            DCK(cudaStreamSynchronize(cs));
            if (d->d_out) cudaFree(d->d_out);
            d->d_out = nullptr; d->out_cap = 0;
            DCK(cudaMalloc((void**)&d->d_out, loff[n_local] * sizeof(float)));
            d->out_cap = loff[n_local];
        }
        while ((int)d->ev_out.size() < chunks) {
            cudaEvent_t e1, e2;
            DCK(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
            DCK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
            d->ev_in.push_back(e1);
            d->ev_out.push_back(e2);
        }
    }
    for (int c = 0; c < chunks; c++) {
        int a, b;
        chunk_range(n_local, chunks, c, a, b);
        if (a >= b) continue;
        std::vector<const float*> L(b - a), R(b - a);
        std::vector<size_t> n(b - a);
        std::vector<float*> o((size_t)(b - a) * pairs * 2);
        for (int k = a; k < b; k++) {
            const int i = mine[k];
            const size_t np = (n_samples[i] + 3) & ~(size_t)3;
            L[k - a] = d->peer_in + in_off[i];
            R[k - a] = d->peer_in + in_off[i] + np;
            n[k - a] = n_samples[i];
            float* base = staged ? d->d_out + loff[k] : d->peer_out + out_off[i];
            for (int sl = 0; sl < pairs * 2; sl++) o[(size_t)(k - a) * pairs * 2 + sl] = base + (size_t)sl * np;
        }
        // the kernels' own loads cross NVLink (PCM from root's memory); so do their stores in the direct form
        if (srt_separate_device(ctx, L.data(), R.data(), n.data(), b - a, unaffected, o.data()))
            return dfail(SRT_ERR_STATE, "rank %d chunk %d: %s", d->rank, c, srt_last_error());
        if (staged) {
            DCK(cudaEventRecord(d->ev_out[c], cs));
            DCK(cudaStreamWaitEvent(d->comm_stream, d->ev_out[c], 0));
            for (int k = a; k < b; k++)
                DCK(cudaMemcpyAsync(d->peer_out + out_off[mine[k]], d->d_out + loff[k], (loff[k + 1] - loff[k]) * sizeof(float), cudaMemcpyDeviceToDevice,
                                    d->comm_stream));
        }
    }
    // completion: a one-word all-reduce behind every rank's kernels (stream order makes their stores visible first)
    DCK(cudaEventRecord(d->ev_done, cs));
    DCK(cudaStreamWaitEvent(d->comm_stream, d->ev_done, 0));
    NCK(ncclAllReduce(d->d_flag, d->d_flag + 1, 1, ncclFloat, ncclSum, d->comm, d->comm_stream));
    return 0;
}
