// callsites.hpp — the reference's verification call sites restated over IVerifier, batch-first.
//
//   voteSet / registerVote            /root/reference/internal/bft/util.go:114-143
//   acceptCommits / acceptPrepares    internal/bft/view.go:144-171
//   processPrepares                   internal/bft/view.go:441-517   (digest match only — prepares are unsigned)
//   processCommits + verifyVote       internal/bft/view.go:519-551, 827-849
//   verifyPrevCommitSignatures        internal/bft/view.go:606-647
//   ValidateLastDecision              internal/bft/viewchanger.go:681-727
//   Controller.HandleRequest          internal/bft/controller.go:233-246
//   Pool.Prune                        internal/bft/requestpool.go:335-354
//
// The reference handles one vote / signature / request per call; here every site that already holds
// a batch hands it to the verifier in one call (IVerifier::*Batch), with identical outcomes.
#pragma once
#include <algorithm>
#include <map>
#include <set>

#include "verifier.hpp"

namespace sbft {

// ---- wire shapes (smartbftprotos/messages.proto:36-54, 92-96) ----
struct PrepareMsg { uint64_t View = 0, Seq = 0; std::string Digest; bool Assist = false; };
struct ProtoSignature { uint64_t Signer = 0; Bytes Value, Msg; };
struct CommitMsg { uint64_t View = 0, Seq = 0; std::string Digest; std::optional<ProtoSignature> Sig; bool Assist = false; };
struct Vote { uint64_t sender = 0; std::optional<PrepareMsg> prepare; std::optional<CommitMsg> commit; };

// ---- voteSet (util.go:114-143) ----
class VoteSet {
  public:
    explicit VoteSet(std::function<bool(uint64_t, const Vote &)> valid) : validVote_(std::move(valid)) {}
    void clear() { voted_.clear(); votes_.clear(); }
    void registerVote(uint64_t voter, const Vote &m) {
        if (!validVote_(voter, m)) return;
        if (voted_.count(voter)) return;  // double vote
        voted_.insert(voter);
        votes_.push_back(m);
        votes_.back().sender = voter;
    }
    const std::vector<Vote> &votes() const { return votes_; }
  private:
    std::function<bool(uint64_t, const Vote &)> validVote_;
    std::set<uint64_t> voted_;
    std::vector<Vote> votes_;
};
inline bool acceptPrepares(uint64_t, const Vote &m) { return m.prepare.has_value(); }  // view.go:146-148
inline bool acceptCommits(uint64_t sender, const Vote &m) {                              // view.go:161-171
    if (!m.commit) return false;
    if (!m.commit->Sig) return false;
    return m.commit->Sig->Signer == sender;
}

struct ViewLog { std::vector<std::string> warnings; };  // the reference tests wait for these lines

// ---- processPrepares (view.go:441-517): Quorum-1 prepares whose digest equals the proposal's ----
// votes arrive in order; returns the voter IDs collected (size Quorum-1) or fewer if the stream ends.
inline std::vector<uint64_t> processPrepares(const Proposal &proposal, int quorum, uint64_t selfID,
                                             const std::vector<Vote> &incoming, ViewLog *log = nullptr) {
    VoteSet prepares(acceptPrepares);
    const std::string expected = proposal.Digest();
    std::vector<uint64_t> voterIDs;
    size_t consumed = 0;
    for (const Vote &m : incoming) {
        if ((int)voterIDs.size() >= quorum - 1) break;
        if (m.sender == selfID) continue;  // view.go:214-217: own messages never reach the vote set
        prepares.registerVote(m.sender, m);
        while (consumed < prepares.votes().size() && (int)voterIDs.size() < quorum - 1) {
            const Vote &v = prepares.votes()[consumed++];
            if (v.prepare->Digest != expected) {
                if (log) log->warnings.push_back("Got wrong digest at processPrepares for prepare with seq " + std::to_string(v.prepare->Seq));
                continue;
            }
            voterIDs.push_back(v.sender);
        }
    }
    return voterIDs;
}

// ---- processCommits (view.go:519-551) with verifyVote (view.go:827-849), batched ----
// All registered votes with a matching digest are verified in ONE VerifyConsenterSigBatch call; the
// first Quorum-1 valid ones in arrival order are returned (the reference's goroutine-per-vote order
// is nondeterministic; any Quorum-1 valid signatures satisfy it).
inline std::vector<Signature> processCommits(IVerifier &verifier, const Proposal &proposal, int quorum, uint64_t selfID,
                                             const std::vector<Vote> &incoming, ViewLog *log = nullptr) {
    VoteSet commits(acceptCommits);
    for (const Vote &m : incoming) {
        if (m.sender == selfID) continue;
        commits.registerVote(m.sender, m);
    }
    const std::string expected = proposal.Digest();
    std::vector<Signature> candidates;
    for (const Vote &v : commits.votes()) {
        if (v.commit->Digest != expected) {  // view.go:829-832
            if (log) log->warnings.push_back("Got wrong digest at processCommits for seq " + std::to_string(v.commit->Seq));
            continue;
        }
        candidates.push_back(Signature{v.commit->Sig->Signer, v.commit->Sig->Value, v.commit->Sig->Msg});
    }
    std::vector<Signature> signatures;
    if (candidates.empty()) return signatures;
    auto res = verifier.VerifyConsenterSigBatch(candidates, proposal);  // view.go:834-838, one call
    for (size_t i = 0; i < candidates.size() && (int)signatures.size() < quorum - 1; i++) {
        if (res[i].second) {  // view.go:839-842
            if (log) log->warnings.push_back("Couldn't verify " + std::to_string(candidates[i].ID) + "'s signature: " + *res[i].second);
            continue;
        }
        signatures.push_back(candidates[i]);
    }
    return signatures;  // decided iff size() == quorum-1; decide() then appends the node's own (view.go:856)
}

// ---- verifyPrevCommitSignatures (view.go:606-647) ----
// Returns (acks, error).  Skipped (empty, nil) when the verification sequence advanced (:616-620).
inline std::pair<std::map<uint64_t, PreparesFrom>, Error> verifyPrevCommitSignatures(IVerifier &verifier, const std::vector<ProtoSignature> &prevCommitSignatures,
                                                                                     const Proposal &prevProp, uint64_t currVerificationSeq) {
    std::map<uint64_t, PreparesFrom> acks;
    if ((uint64_t)prevProp.VerificationSequence != currVerificationSeq) return {acks, std::nullopt};
    std::vector<Signature> sigs;
    for (const auto &s : prevCommitSignatures) sigs.push_back(Signature{s.Signer, s.Value, s.Msg});
    auto res = verifier.VerifyConsenterSigBatch(sigs, prevProp);  // all must verify (:630-638)
    for (size_t i = 0; i < sigs.size(); i++) {
        if (res[i].second) return {{}, Errorf("failed verifying consenter signature of " + std::to_string(sigs[i].ID) + ": " + *res[i].second)};
        PreparesFrom prpf;
        if (!PreparesFrom::Unmarshal(res[i].first, prpf)) return {{}, Errorf("failed unmarshaling auxiliary input from " + std::to_string(sigs[i].ID))};
        acks[sigs[i].ID] = prpf;
    }
    return {acks, std::nullopt};
}

// ---- ValidateLastDecision (viewchanger.go:681-727) ----
struct ViewData {  // messages.proto:65-71 (fields this call site reads)
    uint64_t NextView = 0;
    std::optional<Proposal> LastDecision;  // Metadata empty == nil (genesis)
    std::vector<ProtoSignature> LastDecisionSignatures;
};
inline std::pair<uint64_t, Error> ValidateLastDecision(const ViewData &vd, int quorum, uint64_t n, IVerifier &verifier) {
    (void)n;
    if (!vd.LastDecision) return {0, Errorf("the last decision is not set")};
    if (vd.LastDecision->Metadata.empty()) return {0, std::nullopt};  // genesis proposal
    ViewMetadata md;
    if (!ViewMetadata::Unmarshal(vd.LastDecision->Metadata, md)) return {0, Errorf("unable to unmarshal last decision metadata")};
    if (md.ViewId >= vd.NextView)
        return {0, Errorf("last decision view " + std::to_string(md.ViewId) + " is greater or equal to requested next view " + std::to_string(vd.NextView))};
    int numSigs = (int)vd.LastDecisionSignatures.size();
    if (numSigs < quorum) return {0, Errorf("there are only " + std::to_string(numSigs) + " last decision signatures")};
    std::set<uint64_t> nodes;
    std::vector<Signature> distinct;
    for (const auto &s : vd.LastDecisionSignatures) {
        if (nodes.count(s.Signer)) continue;  // seen signature from this node already
        nodes.insert(s.Signer);
        distinct.push_back(Signature{s.Signer, s.Value, s.Msg});
    }
    auto res = verifier.VerifyConsenterSigBatch(distinct, *vd.LastDecision);  // one call instead of a loop
    int validSig = 0;
    for (size_t i = 0; i < distinct.size(); i++) {
        if (res[i].second) return {0, Errorf("last decision signature is invalid, error: " + *res[i].second)};
        validSig++;
    }
    if (validSig < quorum) return {0, Errorf("there are only " + std::to_string(validSig) + " valid last decision signatures")};
    return {md.LatestSequence, std::nullopt};
}

// ---- request pool fragments (requestpool.go:335-354, controller.go:233-246) ----
class Pool {
  public:
    size_t Size() const { return reqs_.size(); }
    void Submit(const Bytes &req, const RequestInfo &info) { reqs_.push_back({req, info}); }
    // Prune removes requests for which the verifier returns an error — the whole pool in ONE call.
    size_t Prune(IVerifier &verifier) {
        std::vector<Bytes> vec;
        for (auto &r : reqs_) vec.push_back(r.first);
        auto res = verifier.VerifyRequestBatch(vec);
        size_t pruned = 0;
        std::vector<std::pair<Bytes, RequestInfo>> keep;
        for (size_t i = 0; i < reqs_.size(); i++) { if (res[i].second) pruned++; else keep.push_back(reqs_[i]); }
        reqs_.swap(keep);
        return pruned;
    }
    const std::vector<std::pair<Bytes, RequestInfo>> &requests() const { return reqs_; }
  private:
    std::vector<std::pair<Bytes, RequestInfo>> reqs_;
};

// Controller.HandleRequest (controller.go:233-246): not the leader -> dropped without calling the
// verifier; bad request -> not enqueued; good -> Submit.
inline bool HandleRequest(bool iAmTheLeader, IVerifier &verifier, Pool &pool, const Bytes &req, ViewLog *log = nullptr) {
    if (!iAmTheLeader) { if (log) log->warnings.push_back("dropping request: not the leader"); return false; }
    auto res = verifier.VerifyRequest(req);
    if (res.second) { if (log) log->warnings.push_back("Got bad request: " + *res.second); return false; }
    pool.Submit(req, res.first);
    return true;
}

// Controller.MaybePruneRevokedRequests (controller.go:733-746)
inline bool MaybePruneRevokedRequests(uint64_t &cachedVerSeq, IVerifier &verifier, Pool &pool) {
    uint64_t now = verifier.VerificationSequence();
    if (now == cachedVerSeq) return false;
    cachedVerSeq = now;
    pool.Prune(verifier);
    return true;
}

}  // namespace sbft
