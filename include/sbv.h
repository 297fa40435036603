/*
 * sbv.h — C ABI of the B200-native batched signature-verification engine.
 *
 * This is the drop-in boundary behind SmartBFT's application-implemented verifier plug-in:
 *   api.Verifier            /root/reference/pkg/api/dependencies.go:54-71
 *     VerifyConsenterSig    dependencies.go:60-62   (callers: internal/bft/view.go:834-838, 631-635,
 *                                                    internal/bft/viewchanger.go:718)
 *     VerifySignature       dependencies.go:63-64   (callers: viewchanger.go:598, 660, 983, 1022, 1076)
 *     VerifyRequest         dependencies.go:58-59   (callers: internal/bft/controller.go:239, 742-745,
 *                                                    internal/bft/requestpool.go:335-354)
 *     VerifyProposal        dependencies.go:56-57   (caller: view.go:555 — bulk VerifyRequest site)
 *   commit-vote collection  internal/bft/view.go:519-551 (processCommits), 827-849 (verifyVote),
 *                           internal/bft/util.go:114-143 (voteSet), 183-187 (computeQuorum)
 *   digests                 pkg/types/types.go:50-69 (Proposal.Digest), util.go:564-586
 *
 * A cgo (or any FFI) shim binds exactly these symbols; see INTEGRATION.md for the Go side.
 *
 * Conventions
 *  - Return value: 0 on success, < 0 on ENGINE FAULT (CUDA error, out of memory, bad argument).  A
 *    fault is never a verdict: the reference treats `error != nil` from a Verifier as "bad
 *    signature" (view.go:839-842, 386-393), so a host shim must fail-stop on a negative return,
 *    not convert it to a reject.
 *  - Per-item verdicts go to caller-owned arrays: 1 = accept, 0 = reject.
 *  - All pointers are HOST pointers unless the function name ends in _device.  Buffers are only
 *    read during the call and never retained (cgo pointer rules).  Pinned (cudaHostAlloc /
 *    cudaHostRegister) inputs are copied directly; pageable inputs are staged through the
 *    engine's own pinned ring.
 *  - Field elements and scalars are fixed-width big-endian: 32 bytes for P-256, 48 for P-384.
 *  - Accept set = Go crypto/ecdsa (Verify / VerifyASN1): key must be an on-curve affine point with
 *    coordinates < p; r, s in [1, n-1]; e = leftmost min(len, 32|48) digest bytes; R = (e/s)G +
 *    (r/s)Q must not be infinity; accept iff R.x mod n == r.  No low-S rule.
 *  - Thread-safe and re-entrant: the CUDA device is set explicitly per call, so calls may come from any
 *    OS thread (cgo).  Every host-buffer entry point owns a lane (stream, device buffers, pinned staging) for
 *    the duration of the call — six calls proceed concurrently and overlap their copies and kernels; a
 *    seventh waits.  A shard of >= 262,144 items (SBV_CHUNK_ITEMS) is uploaded in chunks on a second stream while
 *    the chunks that have arrived are hashed and verified.  On every return path, faults included, the lane has been drained: no copy into or out of
 *    the caller's buffers is in flight after the call returns, and on a fault the output arrays are untouched
 *    or partially written but never read as verdicts (the caller fail-stops).
 *  - Keys that repeat inside a keys-per-item batch are detected on the device and verified against a per-key
 *    fixed-base table built on the spot (SBV_GROUP_THRESHOLD, default 16 occurrences; 0 disables): the verdicts
 *    are the same bit for bit, only cheaper.
 *  - There is no CPU fallback: without a usable CUDA device sbv_create fails.
 */
#ifndef SBV_H
#define SBV_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sbv_engine sbv_engine;

enum { SBV_P256 = 0, SBV_P384 = 1 };
enum {
    SBV_OK = 0,
    SBV_ERR_ARG = -1,   /* bad argument */
    SBV_ERR_CUDA = -2,  /* CUDA runtime / driver error */
    SBV_ERR_NCCL = -3,  /* NCCL error (multi-device engines only) */
    SBV_ERR_NOMEM = -4
};

/* n_devices in {1,2,4,8}; device_ordinals == NULL means 0..n_devices-1.  With n_devices > 1 the
 * engine shards every batch across the devices and gathers the packed verdict bitmask with NCCL. */
int sbv_create(const int *device_ordinals, int n_devices, sbv_engine **out);
void sbv_destroy(sbv_engine *e);
/* Human-readable description of the last fault on this engine (a thread-local copy: valid until the calling
 * thread's next sbv_last_error). */
const char *sbv_last_error(const sbv_engine *e);
int sbv_device_count(const sbv_engine *e);

/* ECDSA verify, digests supplied.  SoA arrays of n fixed-width big-endian values. */
int sbv_verify_batch(sbv_engine *e, uint8_t curve, size_t n, const uint8_t *r, const uint8_t *s,
                     const uint8_t *qx, const uint8_t *qy, const uint8_t *digest, uint8_t digest_len,
                     uint8_t *ok);

/* Same, with every input already resident on device `device_index` (0-based position in the
 * engine's device list) and the verdicts left on the device.  Enqueued on `cuda_stream`
 * (a cudaStream_t; NULL = the legacy default stream) and NOT synchronised. */
int sbv_verify_batch_device(sbv_engine *e, int device_index, uint8_t curve, size_t n, const uint8_t *d_r,
                            const uint8_t *d_s, const uint8_t *d_qx, const uint8_t *d_qy,
                            const uint8_t *d_digest, uint8_t digest_len, uint8_t *d_ok, void *cuda_stream);

/* DER front end: sigs = concatenated ASN.1 SEQUENCE{INTEGER r, INTEGER s}, sig_off[n+1].  Parsed
 * with crypto/ecdsa.VerifyASN1 strictness (minimal lengths, minimal non-negative integers, no
 * trailing bytes); malformed items reject.  qxy = n * (X||Y). */
int sbv_verify_batch_der(sbv_engine *e, uint8_t curve, size_t n, const uint8_t *sigs, const uint32_t *sig_off,
                         const uint8_t *qxy, const uint8_t *digest, uint8_t digest_len, uint8_t *ok);

/* SHA-256 over a ragged batch: msgs concatenated, msg_off[n+1] byte offsets.  digest_out = 32n. */
int sbv_sha256_batch(sbv_engine *e, size_t n, const uint8_t *msgs, const uint64_t *msg_off, uint8_t *digest_out);

/* Fused SHA-256 -> ECDSA verify (VerifyRequest / VerifySignature shape): the message digest never
 * leaves the device.  digest_out may be NULL. */
int sbv_hash_verify_batch(sbv_engine *e, uint8_t curve, size_t n, const uint8_t *msgs, const uint64_t *msg_off,
                          const uint8_t *r, const uint8_t *s, const uint8_t *qx, const uint8_t *qy,
                          uint8_t *digest_out, uint8_t *ok);

/* Mixed-curve batch: curve_tag[i] in {SBV_P256, SBV_P384}; every field is stored in a 48-byte
 * slot (P-256 values right-aligned, i.e. 16 leading zero bytes); digest is 32 bytes per item. */
int sbv_verify_mixed(sbv_engine *e, size_t n, const uint8_t *curve_tag, const uint8_t *r48, const uint8_t *s48,
                     const uint8_t *qx48, const uint8_t *qy48, const uint8_t *digest32, uint8_t *ok);

/* Distinct-signer quorum count per consensus instance (processCommits, view.go:519-551).
 * Votes are given in arrival order.  A vote is registered iff signer == sender and sender !=
 * self_id[instance] and the sender has no earlier registered vote in the instance
 * (view.go:161-171, util.go:130-143); a registered vote is valid iff digest_match && ok
 * (view.go:829-842).  valid_count[i] = number of valid votes; reached[i] = valid_count[i] >=
 * threshold (the caller passes Quorum-1, view.go:531).  self_id may be NULL (no self filter). */
int sbv_quorum(sbv_engine *e, size_t n_votes, const uint32_t *instance, const uint16_t *sender,
               const uint16_t *signer, const uint8_t *digest_match, const uint8_t *ok, size_t n_instances,
               const uint16_t *self_id, uint32_t threshold, uint32_t *valid_count, uint8_t *reached);

/* Prepare collection in batch form (View.processPrepares, view.go:441-517): prepares carry no signature; the
 * first prepare of a sender burns its slot (util.go:130-143) and counts iff digest_match (view.go:452-459);
 * reached[i] = match_count[i] >= threshold (the caller passes Quorum-1, view.go:446). */
int sbv_prepare_quorum(sbv_engine *e, size_t n_votes, const uint32_t *instance, const uint16_t *sender,
                       const uint8_t *digest_match, size_t n_instances, const uint16_t *self_id, uint32_t threshold,
                       uint32_t *match_count, uint8_t *reached);

/* Commit-vote verification AND quorum collection in one call (verifyVote + processCommits, view.go:519-551,
 * 827-849; BASELINE configs[3]).  The n_votes signatures (SoA as in sbv_verify_batch) are verified, the verdicts stay
 * on the device and feed the distinct-signer count of sbv_quorum.  Votes must be grouped by instance with
 * non-decreasing instance ids, in arrival order inside an instance.  A multi-device engine shards BY INSTANCE so that
 * every count is local; the packed verdict mask and the packed `reached` mask travel in one NCCL all-gather.
 * Outputs: ok[n_votes], valid_count[n_instances], reached[n_instances]. */
int sbv_verify_quorum(sbv_engine *e, uint8_t curve, size_t n_votes, const uint8_t *r, const uint8_t *s, const uint8_t *qx,
                      const uint8_t *qy, const uint8_t *digest, uint8_t digest_len, const uint32_t *instance,
                      const uint16_t *sender, const uint16_t *signer, const uint8_t *digest_match, size_t n_instances,
                      const uint16_t *self_id, uint32_t threshold, uint8_t *ok, uint32_t *valid_count, uint8_t *reached);

/* computeQuorum(n) -> (q, f), internal/bft/util.go:183-187. */
void sbv_compute_quorum(uint64_t n, uint32_t *q, uint32_t *f);

/* Consenter key registry.  Keys are configuration in the reference: they change only with a
 * reconfiguration, i.e. a new VerificationSequence (dependencies.go:65-66).  sbv_set_keys replaces
 * the registry and precomputes, on every device, a fixed-base comb table per key
 * (8-bit signed windows: 33 x 128 affine points = 264 KiB per P-256 key, a few milliseconds for a thousand keys);
 * slot i of the registry is key i of this call.
 * xy = n * 96 bytes: X and Y in 48-byte slots (P-256 values right-aligned). */
int sbv_set_keys(sbv_engine *e, uint64_t verification_seq, size_t n, const uint64_t *ids, const uint8_t *curve,
                 const uint8_t *xy);

/* ECDSA verify against REGISTERED keys: key_slot[i] indexes the registry of sbv_set_keys.  Same accept
 * set as sbv_verify_batch; an unknown slot, a slot of another curve or an invalid registered key
 * rejects.  Both scalar multiplications are fixed-base (no doublings), which is ~5x less work than
 * the keys-per-item entry point. */
int sbv_verify_registered(sbv_engine *e, uint8_t curve, size_t n, const uint32_t *key_slot, const uint8_t *r,
                          const uint8_t *s, const uint8_t *digest, uint8_t digest_len, uint8_t *ok);
/* Fused SHA-256 -> registered-key verify (VerifyConsenterSig / VerifySignature / VerifyRequest with
 * registered consenter or client keys): messages hashed on the device. */
int sbv_hash_verify_registered(sbv_engine *e, uint8_t curve, size_t n, const uint8_t *msgs, const uint64_t *msg_off,
                               const uint32_t *key_slot, const uint8_t *r, const uint8_t *s, uint8_t *ok);
int sbv_verify_registered_device(sbv_engine *e, int device_index, uint8_t curve, size_t n, const uint32_t *d_key_slot,
                                 const uint8_t *d_r, const uint8_t *d_s, const uint8_t *d_digest, uint8_t digest_len,
                                 uint8_t *d_ok, void *cuda_stream);

/* ---- one process per GPU (a Go host may run one node process per device; bench.py does under torchrun) ----
 * The engine of every process is one RANK; the only exchange is the all-gather of packed verdict / quorum bitmasks
 * over NCCL (NVLink / NVSwitch).  Rank 0 calls sbv_comm_unique_id and ships the 128 bytes to the others (any side
 * channel); every rank then calls sbv_comm_init_rank, which adds one CHANNEL (communicator) and returns its index.
 * The collectives of a channel must be issued in the same order on every rank, so concurrent caller threads take
 * one channel each (create as many as there are threads, in the same order on every rank). */
int sbv_comm_unique_id(uint8_t *id128);
int sbv_comm_init_rank(sbv_engine *e, const uint8_t *id128, int nranks, int rank);
int sbv_comm_ranks(const sbv_engine *e);
/* Device form: packs n verdict bytes into a bitmask and all-gathers the masks of all ranks,
 * d_mask_all[rank * ceil(n/32) + w]; enqueued on cuda_stream, not synchronised.  Every rank passes the same n. */
int sbv_gather_verdicts_device(sbv_engine *e, int channel, const uint8_t *d_ok, size_t n, uint32_t *d_mask_all, void *cuda_stream);
/* All-gather of `words` 32-bit words per rank, in place: the sender's words sit at d_all + rank * words. */
int sbv_gather_words_device(sbv_engine *e, int channel, uint32_t *d_all, size_t words, void *cuda_stream);
/* Host form: sbv_verify_batch for this rank's n items + the gather: ok[n] = this rank's verdict bytes,
 * mask_all[nranks * ceil(n/32)] = the packed verdicts of every rank. */
int sbv_verify_batch_ranked(sbv_engine *e, int channel, uint8_t curve, size_t n, const uint8_t *r, const uint8_t *s,
                            const uint8_t *qx, const uint8_t *qy, const uint8_t *digest, uint8_t digest_len, uint8_t *ok,
                            uint32_t *mask_all);

/* Pinned (page-locked, portable) host memory for batches the host marshals itself: buffers from here are DMA'd
 * directly by every entry point (no staging copy).  A cgo shim keeps C memory anyway (cgo pointer rules), so its
 * batch buffers should come from here.  NULL on failure. */
void *sbv_host_alloc(size_t bytes);
void sbv_host_free(void *p);

/* Introspection for benchmarks: number of kernel launches issued by this engine so far. */
uint64_t sbv_kernel_launches(const sbv_engine *e);
/* Optional CUDA-event timing inside every verify launch (off by default).  sbv_profile_read sums, over all
 * devices, prep_ms = launch start .. end of the scalar preparation (includes the key grouping) and verify_ms = the
 * verification kernels alone (the two halves of the fixed-base verification when keys were grouped, the generic kernel
 * otherwise), and
 * resets; the caller synchronises the streams it used first. */
int sbv_profile_enable(sbv_engine *e, int on);
int sbv_profile_read(sbv_engine *e, double *prep_ms, double *verify_ms, uint64_t *n_launch_pairs);
/* Peak-rate probe: dependent-free IMAD.WIDE.U32 loop on device 0; returns MAC32/s (0 on fault). */
double sbv_probe_mad_rate(sbv_engine *e);

#ifdef __cplusplus
}
#endif
#endif /* SBV_H */
