// quorum.cuh — distinct-signer valid-vote counting per consensus instance.
//
// Restates, for a batch of instances, what View.processCommits does one vote at a time
// (/root/reference/internal/bft/view.go:519-551):
//   * a commit vote is registered only if Signature.Signer == sender     (view.go:161-171)
//   * the view never registers its own vote                              (sender == self)
//   * one registered vote per sender — the first one wins, later ones are dropped even if the
//     first turns out invalid                                            (util.go:130-143)
//   * a registered vote is valid iff its digest matches the proposal's and VerifyConsenterSig
//     accepted it                                                        (view.go:829-842)
//   * the instance is decided once Quorum-1 valid foreign votes exist    (view.go:531)
// Votes of one instance must be contiguous and in arrival order.  One thread per vote scans back
// over its own instance's earlier votes (<= N-1 of them), so no sort and no shared state.
#pragma once
#include <stdint.h>

namespace sbv {

__global__ void k_quorum_count(uint32_t n_votes, const uint32_t *__restrict__ instance, const uint16_t *__restrict__ sender,
                               const uint16_t *__restrict__ signer, const uint8_t *__restrict__ digest_match,
                               const uint8_t *__restrict__ ok, const uint16_t *__restrict__ self_id,
                               uint32_t inst_base, uint32_t n_instances, uint32_t *__restrict__ valid_count) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_votes) return;
    const uint32_t inst = instance[v] - inst_base;  // a device holds the instances [inst_base, inst_base + n_instances)
    if (inst >= n_instances) return;
    const uint16_t snd = sender[v];
    if (signer[v] != snd) return;
    if (self_id && self_id[inst] == snd) return;
    if (!(digest_match[v] && (!ok || ok[v]))) return;  // registered or not, it cannot count (ok == NULL: prepares carry no signature)
    for (uint32_t j = v; j-- > 0;) {
        if (instance[j] - inst_base != inst) break;
        if (sender[j] == snd && signer[j] == snd) return;  // an earlier registered vote burnt the slot
    }
    atomicAdd(valid_count + inst, 1u);
}

__global__ void k_quorum_reached(uint32_t n_instances, const uint32_t *__restrict__ valid_count, uint32_t threshold,
                                 uint8_t *__restrict__ reached) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_instances) reached[i] = valid_count[i] >= threshold ? 1 : 0;
}

// ok bytes -> packed bitmask (bit i of word i/32), for the cross-device gather
__global__ void k_pack_bits(uint32_t n, const uint8_t *__restrict__ ok, uint32_t *__restrict__ mask) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t bit = (i < n && ok[i]) ? 1u : 0u;
    const uint32_t word = __ballot_sync(0xffffffffu, bit);
    if ((threadIdx.x & 31) == 0 && i < n) mask[i >> 5] = word;
}

}  // namespace sbv
