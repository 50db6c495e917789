/*
 * srt_dispatch.h — the stream dispatcher for a multi-GPU box, C ABI of libspleeterrt_dispatch.so (links libnccl + libspleeterrt_b200).
 *
 * The hot path shards by STREAM: units (stream, tile, stem) share nothing, so there is no collective between
 * kernels.  NCCL (NVLink 5 / NVSwitch on the B200 box) is used only to move whole streams: one process per GPU, every rank
 * holds a full weight replica (one ncclBroadcast at start-up), and for a batch that lives on ONE rank ("root": the process that
 * decoded or received the audio) the dispatcher
 *     scatters   stream i -> rank i mod world          grouped ncclSend / ncclRecv of the PCM, device to device
 *     separates  every rank runs srt_separate_device on its share (root included)
 *     gathers    the stems back into root's buffers     grouped ncclSend / ncclRecv, straight into the caller's pointers
 * in `chunks` pipelined groups, so a rank computes chunk c while chunk c+1 arrives and chunk c-1 leaves.  Nothing passes
 * through host memory.  The reference has no counterpart (one file per process, Executable/main.c): this is BASELINE.json
 * configs[3] ("batch = 1024 streams sharded 8 x B200 via NCCL stream dispatch").
 *
 * Every rank calls every function in the same order with the same (world, root, n_streams, n_samples, chunks).  Returns 0 or a
 * negative srt_status; srt_dispatch_last_error() has the text.
 */
#ifndef SRT_DISPATCH_H
#define SRT_DISPATCH_H
#include <stddef.h>

#include "srt_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

#define SRT_DISPATCH_ID_BYTES 128 /* sizeof(ncclUniqueId) */

typedef struct srt_dispatch srt_dispatch;

/* rank 0: create the rendezvous token; the launcher hands the same bytes to every rank (a file, a socket, MPI, torch.distributed) */
int srt_dispatch_get_id(unsigned char id[SRT_DISPATCH_ID_BYTES]);
/* ncclCommInitRank on `device`; collective over all ranks */
int srt_dispatch_create(const unsigned char id[SRT_DISPATCH_ID_BYTES], int world, int rank, int device, srt_dispatch** out);
void srt_dispatch_destroy(srt_dispatch* d);
const char* srt_dispatch_last_error(void);

/* Weights once: `coeffs` = n_nets x SRT_COEFF_FLOATS floats of HOST memory on every rank, filled on `root`; on return every rank
 * holds root's bytes (staged through device memory, one ncclBroadcast). */
int srt_dispatch_broadcast_weights(srt_dispatch* d, float* coeffs, int n_nets, int root);
/* small host arrays the same way (stream lengths): n 64-bit values */
int srt_dispatch_broadcast_sizes(srt_dispatch* d, size_t* values, int n, int root);

/* One batch.  On root: d_pcmL[i], d_pcmR[i] = device pointers (root's GPU) to stream i's channels, n_samples[i] floats each, for
 * ALL n_streams streams; d_out[(i * pairs + q) * 2 + c] = device pointers for the results (pairs = srt_output_pairs(ctx)).  Other
 * ranks pass NULL for the three pointer arrays.  n_samples / n_streams / unaffected / chunks are the same on all ranks.
 * ctx = this rank's context (srt_create with max_batch_images >= the tiles of its largest chunk).  Enqueues everything and
 * returns; srt_dispatch_wait() blocks until this rank's part (on root: all results) is complete. */
int srt_dispatch_separate_device(srt_dispatch* d, srt_ctx* ctx, int root, const float* const* d_pcmL, const float* const* d_pcmR,
                                 const size_t* n_samples, int n_streams, const float* unaffected, float* const* d_out, int chunks);
int srt_dispatch_wait(srt_dispatch* d);
/* The CUDA stream (cudaStream_t) the transfers of the last call ran on, for callers that time with CUDA events. */
void* srt_dispatch_comm_stream(srt_dispatch* d);

/* ---- peer-memory mode: no copies at all ---------------------------------------------------------------------------------
 * On an NVSwitch box every GPU can load and store every other GPU's memory at NVLink speed.  Root allocates the batch's PCM and
 * stem buffers ONCE (srt_dispatch_peer_buffers), the other ranks map them (CUDA IPC handles travel over the NCCL communicator),
 * and every rank then runs its share of the streams with its STFT kernel LOADING the PCM straight from root's memory and its
 * overlap-add kernel STORING the stems straight into root's memory: the transfer is the kernels' own global accesses, spread
 * over the whole step, instead of a send/recv phase before and after.  NCCL only carries the handles and a one-word
 * all-reduce that tells root when everybody is done.
 *   layout of root's buffers (srt_dispatch_peer_layout, host only): stream i's channels at in_off[i] (L) and in_off[i] + np_i (R),
 *   its output pair q / channel c at out_off[i] + (q * 2 + c) * np_i, np_i = n_samples[i] rounded up to 4 floats.
 * srt_dispatch_peer_buffers: collective; in_floats / out_floats = capacities (root's values are used); returns root's buffers
 *   as seen from this rank (on root: its own allocation).  The buffers live until the dispatcher is destroyed or re-requested.
 * srt_dispatch_separate_peer: collective; every rank separates streams rank, rank + world, ... in `chunks` calls of
 *   srt_separate_device.  chunks == 1: the overlap-add kernel stores into root's memory directly.  chunks > 1: a chunk's stems are
 *   written locally and leave with one copy-engine transfer per stream (cudaMemcpyAsync into the mapping) while the next chunk
 *   computes - on 8 GPUs the direct form ends in a burst of stores from seven GPUs into one (they run in lockstep), the
 *   pipelined form hides all but the last chunk's transfer.  srt_dispatch_wait() on root returns when all stems are in the buffer. */
int srt_dispatch_peer_layout(const size_t* n_samples, int n_streams, int pairs, size_t* in_off, size_t* out_off,
                             size_t* in_floats, size_t* out_floats);
int srt_dispatch_peer_buffers(srt_dispatch* d, int root, size_t in_floats, size_t out_floats, float** d_in, float** d_out);
int srt_dispatch_separate_peer(srt_dispatch* d, srt_ctx* ctx, int root, const size_t* n_samples, int n_streams,
                               const float* unaffected, int chunks);

/* ---- the schedule, without NCCL (host only; unit-tested on the CPU) ----------------------------------------------
 * The point-to-point operations rank `rank` issues for one batch, in issue order, as rows of 6 ints:
 *   {group, kind (0 = send, 1 = recv), peer, stream, slot, count}
 * group: operations of one ncclGroupStart/End; slot: 0 = L, 1 = R (scatter) or 2 + q * 2 + c (gather of output pair q, channel c).
 * Two ranks' lists pair up send for recv in the same order for every (sender, receiver) - the property NCCL needs.
 * Returns the number of rows (written to `rows` if it holds at least that many, cap_rows in rows), or a negative status. */
long long srt_dispatch_schedule(int world, int rank, int root, const size_t* n_samples, int n_streams, int pairs, int chunks,
                                int* rows, long long cap_rows);
/* streams of `rank`, in local order; chunk c of the rank = local indices [c * per, (c + 1) * per), per = ceil(n_local / chunks) */
int srt_dispatch_local_streams(int world, int rank, int n_streams, int* ids /* may be NULL */);

#ifdef __cplusplus
}
#endif
#endif
