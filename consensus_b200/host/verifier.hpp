// verifier.hpp — C++ mirror of the reference's verifier plug-in and its GPU implementation.
//
//   IVerifier      = api.Verifier              /root/reference/pkg/api/dependencies.go:54-71
//                    (+ batch forms of VerifyConsenterSig / VerifyRequest; the defaults loop over the
//                    single forms, so a mock only implements the seven reference methods)
//   GpuVerifier    = the application-side implementation on libsbv.so (include/sbv.h), using the
//                    signed-bytes convention of INTEGRATION.md:
//                        Msg   = SHA-256(asn1(Proposal)) (32 raw bytes) || aux
//                        Value = DER(r, s) over SHA-256(Msg), key registered for Signature.ID
//   Aggregator     = deadline-flush coalescing of concurrent single-signature calls
//                    (view.go:537-541 spawns one goroutine per commit vote)
//
// Engine faults never become verdicts: they throw EngineFault, which the embedding
// application must treat as fatal — exactly as the reference panics on unrecoverable local errors.
#pragma once
#include <array>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <utility>

#include "../../include/sbv.h"
#include "sbft.hpp"

namespace sbft {

struct EngineFault : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ------------------------------------------------------------------------------------------------
struct IVerifier {
    virtual ~IVerifier() = default;
    virtual std::pair<std::vector<RequestInfo>, Error> VerifyProposal(const Proposal &proposal) = 0;          // :56-57
    virtual std::pair<RequestInfo, Error> VerifyRequest(const Bytes &val) = 0;                                // :58-59
    virtual std::pair<Bytes, Error> VerifyConsenterSig(const Signature &signature, const Proposal &prop) = 0; // :60-62
    virtual Error VerifySignature(const Signature &signature) = 0;                                            // :63-64
    virtual uint64_t VerificationSequence() = 0;                                                              // :65-66
    virtual std::vector<RequestInfo> RequestsFromProposal(const Proposal &proposal) = 0;                      // :67-68
    virtual Bytes AuxiliaryData(const Bytes &msg) = 0;                                                        // :69-70

    // Batch forms for the call sites that already hold a batch (view.go:630-644,
    // viewchanger.go:702-722, requestpool.go:339-351).  Same per-item result as the single form.
    virtual std::vector<std::pair<Bytes, Error>> VerifyConsenterSigBatch(const std::vector<Signature> &sigs, const Proposal &prop) {
        std::vector<std::pair<Bytes, Error>> out;
        for (const auto &s : sigs) out.push_back(VerifyConsenterSig(s, prop));
        return out;
    }
    virtual std::vector<std::pair<RequestInfo, Error>> VerifyRequestBatch(const std::vector<Bytes> &reqs) {
        std::vector<std::pair<RequestInfo, Error>> out;
        for (const auto &r : reqs) out.push_back(VerifyRequest(r));
        return out;
    }
};

// ------------------------------------------------------------------------------------------------
// Strict DER SEQUENCE{INTEGER r, INTEGER s} (crypto/ecdsa.VerifyASN1 rules) -> 32-byte r, s.
inline bool parse_der_sig(const Bytes &sig, uint8_t r[32], uint8_t s[32]) {
    auto rd_int = [](const uint8_t *&p, const uint8_t *end, uint8_t *out) {
        if (end - p < 2 || p[0] != 0x02) return false;
        size_t len = p[1];
        p += 2;
        if ((len & 0x80) || len == 0 || (size_t)(end - p) < len) return false;
        if (p[0] & 0x80) return false;
        if (len > 1 && p[0] == 0 && !(p[1] & 0x80)) return false;
        const uint8_t *v = p; size_t vl = len;
        if (vl > 1 && v[0] == 0) { v++; vl--; }
        if (vl > 32) return false;
        memset(out, 0, 32); memcpy(out + 32 - vl, v, vl);
        p += len;
        return true;
    };
    const uint8_t *p = sig.data(), *end = p + sig.size();
    if (sig.size() < 2 || p[0] != 0x30) return false;
    size_t len;
    if (p[1] < 0x80) { len = p[1]; p += 2; }
    else if (p[1] == 0x81) { if (sig.size() < 3 || p[2] < 0x80) return false; len = p[2]; p += 3; }
    else return false;
    if ((size_t)(end - p) != len) return false;
    return rd_int(p, end, r) && rd_int(p, end, s) && p == end;
}

// ---- pinned host memory from the engine (sbv_host_alloc): what the engine DMAs from without a staging copy ----
struct PinnedBuf {
    uint8_t *p = nullptr;
    size_t cap = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    ~PinnedBuf() { if (p) sbv_host_free(p); }
    void reserve(size_t n, size_t keep = 0) {
        if (n <= cap) return;
        size_t nc = std::max(n, cap * 2 + 4096);
        uint8_t *q = (uint8_t *)sbv_host_alloc(nc);
        if (!q) throw EngineFault("sbv_host_alloc failed");
        if (p && keep) memcpy(q, p, keep);
        if (p) sbv_host_free(p);
        p = q; cap = nc;
    }
};

// One unit of work for the engine: verify (r, s) by key (X||Y) over SHA-256(message).
struct SigItem {
    uint8_t r[32], s[32];
    uint32_t slot = 0;  // index into the engine's key registry (sbv_set_keys)
    Bytes message;
};
// Verifies a batch; returns one verdict byte per item.  Production: GpuVerifier::engine_batch.
using BatchFn = std::function<std::vector<uint8_t>(const std::vector<SigItem> &)>;

// ------------------------------------------------------------------------------------------------
class Aggregator {
  public:
    Aggregator(BatchFn fn, std::chrono::microseconds window, size_t max_batch)
        : fn_(std::move(fn)), window_(window), max_(max_batch), open_(std::make_shared<Batch>()), th_([this] { run(); }) {}
    ~Aggregator() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        th_.join();
        std::unique_lock<std::mutex> lk(mu_);
        idle_cv_.wait(lk, [&] { return inflight_ == 0; });
    }
    // Blocks until the batch this item joined has been verified; returns the item's verdict.
    bool submit(SigItem item) {
        std::unique_lock<std::mutex> lk(mu_);
        auto b = open_;
        size_t idx = b->items.size();
        b->items.push_back(std::move(item));
        if (idx == 0) { b->deadline = std::chrono::steady_clock::now() + window_; cv_.notify_all(); }
        if (b->items.size() >= max_) flush_locked(lk);
        b->cv.wait(lk, [&] { return b->done; });
        if (b->fault) throw EngineFault(*b->fault);
        return b->ok[idx] != 0;
    }
    uint64_t batches() const { return batches_; }
    uint64_t items() const { return items_; }

  private:
    struct Batch {
        std::vector<SigItem> items;
        std::vector<uint8_t> ok;
        std::chrono::steady_clock::time_point deadline;
        bool done = false;
        std::optional<std::string> fault;
        std::condition_variable cv;
    };
    // Closes the open batch and runs it on its own thread: neither the deadline thread nor the caller that filled the
    // batch waits for the engine, so consecutive batches overlap on the engine's lanes.
    void flush_locked(std::unique_lock<std::mutex> &) {
        auto b = open_;
        if (b->items.empty()) return;
        open_ = std::make_shared<Batch>();
        inflight_++;
        std::thread([this, b] {
            try { b->ok = fn_(b->items); } catch (const std::exception &ex) { b->fault = ex.what(); }
            std::lock_guard<std::mutex> lk(mu_);
            batches_++; items_ += b->items.size();
            b->done = true;
            b->cv.notify_all();
            inflight_--;
            idle_cv_.notify_all();
        }).detach();
    }
    void run() {
        std::unique_lock<std::mutex> lk(mu_);
        while (!stop_) {
            if (open_->items.empty()) { cv_.wait(lk, [&] { return stop_ || !open_->items.empty(); }); continue; }
            auto dl = open_->deadline;
            if (cv_.wait_until(lk, dl, [&] { return stop_; })) break;
            if (!open_->items.empty() && std::chrono::steady_clock::now() >= open_->deadline) flush_locked(lk);
        }
        flush_locked(lk);
    }
    BatchFn fn_;
    std::chrono::microseconds window_;
    size_t max_;
    std::mutex mu_;
    std::condition_variable cv_, idle_cv_;
    std::shared_ptr<Batch> open_;
    bool stop_ = false;
    int inflight_ = 0;
    uint64_t batches_ = 0, items_ = 0;
    std::thread th_;
};

// ------------------------------------------------------------------------------------------------
// Request framing used by GpuVerifier (application-defined in the reference, node.go:250-265):
//   request := u16be siglen || sig(DER) || u32be clen || client || u32be ilen || id || payload
//   signed  := request[2 + siglen :]
struct ParsedRequest {
    Bytes sig, signedBytes;
    std::string client, id;
};
inline bool parse_request(const Bytes &req, ParsedRequest &out) {
    if (req.size() < 2) return false;
    size_t sl = (size_t)req[0] << 8 | req[1], p = 2;
    if (req.size() < p + sl + 8) return false;
    out.sig.assign(req.begin() + p, req.begin() + p + sl);
    p += sl;
    out.signedBytes.assign(req.begin() + p, req.end());
    auto rd = [&](std::string &s) {
        if (req.size() < p + 4) return false;
        size_t l = (size_t)req[p] << 24 | (size_t)req[p + 1] << 16 | (size_t)req[p + 2] << 8 | req[p + 3];
        p += 4;
        if (req.size() < p + l) return false;
        s.assign(req.begin() + p, req.begin() + p + l);
        p += l;
        return true;
    };
    return rd(out.client) && rd(out.id);
}
inline Bytes frame_request(const Bytes &sig, const std::string &client, const std::string &id, const Bytes &payload) {
    Bytes o{(uint8_t)(sig.size() >> 8), (uint8_t)sig.size()};
    o.insert(o.end(), sig.begin(), sig.end());
    auto wr = [&](const std::string &s) { for (int k = 3; k >= 0; k--) o.push_back((uint8_t)(s.size() >> (8 * k))); o.insert(o.end(), s.begin(), s.end()); };
    wr(client); wr(id);
    o.insert(o.end(), payload.begin(), payload.end());
    return o;
}
inline Bytes signed_part(const std::string &client, const std::string &id, const Bytes &payload) {
    Bytes full = frame_request({}, client, id, payload);
    return Bytes(full.begin() + 2, full.end());
}
// proposal payload := repeated (u32be len || request)
inline bool split_requests(const Bytes &payload, std::vector<Bytes> &out) {
    size_t p = 0;
    while (p < payload.size()) {
        if (payload.size() < p + 4) return false;
        size_t l = (size_t)payload[p] << 24 | (size_t)payload[p + 1] << 16 | (size_t)payload[p + 2] << 8 | payload[p + 3];
        p += 4;
        if (payload.size() < p + l) return false;
        out.emplace_back(payload.begin() + p, payload.begin() + p + l);
        p += l;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
class GpuVerifier : public IVerifier {
  public:
    // devices: CUDA ordinals (1, 2, 4 or 8 of one box).  Throws EngineFault without a usable GPU.
    explicit GpuVerifier(const std::vector<int> &devices, std::chrono::microseconds window = std::chrono::microseconds(200),
                         size_t max_batch = 65536) {
        if (sbv_create(devices.data(), (int)devices.size(), &eng_) != SBV_OK) throw EngineFault("sbv_create failed (no CPU fallback)");
        agg_ = std::make_unique<Aggregator>([this](const std::vector<SigItem> &v) { return engine_batch(v); }, window, max_batch);
    }
    ~GpuVerifier() override { agg_.reset(); sbv_destroy(eng_); }

    // Keys are registered with the engine (sbv_set_keys builds a fixed-base comb table per key), so every
    // verification below is the registered-key path.  The registry is rebuilt lazily after changes —
    // keys change only with a reconfiguration (a new verification sequence).
    void SetConsenterKey(uint64_t id, const uint8_t xy[64]) { std::lock_guard<std::mutex> lk(mu_); consenters_[id] = slot_of(xy); }
    void SetClientKey(const std::string &client, const uint8_t xy[64]) { std::lock_guard<std::mutex> lk(mu_); clients_[client] = slot_of(xy); }
    void SetVerificationSequence(uint64_t v) { std::lock_guard<std::mutex> lk(mu_); verSeq_ = v; }
    // Keys change only with a reconfiguration (a new verification sequence, dependencies.go:65-66): drop the old
    // registry before registering the new configuration's keys, so rotated keys do not pile up in HBM.
    void ResetKeys() {
        std::lock_guard<std::mutex> lk(mu_);
        registry_.clear(); slots_.clear(); consenters_.clear(); clients_.clear();
        dirty_ = true;
    }
    uint32_t ConsenterSlot(uint64_t id) { std::lock_guard<std::mutex> lk(mu_); auto k = consenters_.find(id); return k == consenters_.end() ? 0xffffffffu : k->second; }
    Aggregator &aggregator() { return *agg_; }
    sbv_engine *engine() { return eng_; }

    // One engine call: SHA-256 of every message and ECDSA-P256 verification against the registered
    // keys, both on the GPU.
    std::vector<uint8_t> engine_batch(const std::vector<SigItem> &items) {
        sync_registry();  // also with an empty batch: callers use that to push the registry to the engine
        const size_t n = items.size();
        std::vector<uint8_t> ok(n);
        if (n == 0) return ok;
        // flatten straight into pinned memory (one block per calling thread, reused): r | s | slot | off | msgs
        static thread_local PinnedBuf pin;
        size_t total = 0;
        for (const auto &it : items) total += it.message.size();
        const size_t o_s = 32 * n, o_slot = 64 * n, o_off = o_slot + 4 * n + (8 - (4 * n) % 8) % 8, o_msgs = o_off + 8 * (n + 1);
        pin.reserve(o_msgs + total + 16);
        uint8_t *r = pin.p, *s = pin.p + o_s, *msgs = pin.p + o_msgs;
        uint32_t *slot = (uint32_t *)(pin.p + o_slot);
        uint64_t *off = (uint64_t *)(pin.p + o_off);
        size_t pos = 0;
        off[0] = 0;
        for (size_t i = 0; i < n; i++) {
            memcpy(r + 32 * i, items[i].r, 32); memcpy(s + 32 * i, items[i].s, 32);
            slot[i] = items[i].slot;
            if (!items[i].message.empty()) memcpy(msgs + pos, items[i].message.data(), items[i].message.size());
            pos += items[i].message.size();
            off[i + 1] = pos;
        }
        int rc = sbv_hash_verify_registered(eng_, SBV_P256, n, msgs, off, slot, r, s, ok.data());
        if (rc != SBV_OK) throw EngineFault(std::string("sbv_hash_verify_registered: ") + sbv_last_error(eng_));
        return ok;
    }

    // CommitSignaturesDigest for MANY signature sets (internal/bft/util.go:564-595): DER framing on the host, the
    // SHA-256 chains on the GPU in one call.  An empty set yields an empty digest (nil, util.go:565-567).
    std::vector<Bytes> CommitSignaturesDigestBatch(const std::vector<std::vector<Signature>> &sets) {
        std::vector<Bytes> out(sets.size());
        Bytes blob;
        std::vector<uint64_t> off{0};
        std::vector<size_t> where;
        for (size_t i = 0; i < sets.size(); i++) {
            if (sets[i].empty()) continue;
            Bytes d = CommitSignaturesDer(sets[i]);
            blob.insert(blob.end(), d.begin(), d.end());
            off.push_back(blob.size());
            where.push_back(i);
        }
        if (where.empty()) return out;
        Bytes dig(where.size() * 32);
        if (sbv_sha256_batch(eng_, where.size(), blob.data(), off.data(), dig.data()) != SBV_OK)
            throw EngineFault(std::string("sbv_sha256_batch: ") + sbv_last_error(eng_));
        for (size_t k = 0; k < where.size(); k++) out[where[k]] = Bytes(dig.begin() + 32 * k, dig.begin() + 32 * k + 32);
        return out;
    }

    // Proposal.Digest for MANY proposals (pkg/types/types.go:50-69): DER framing on the host, every SHA-256
    // chain on the GPU in one sbv_sha256_batch call.  (One digest is a single sequential chain and stays
    // on the host: Proposal::Digest.)  Returns the hex strings the reference compares in verifyVote.
    std::vector<std::string> DigestBatch(const std::vector<Proposal> &props) {
        std::vector<std::string> out(props.size());
        if (props.empty()) return out;
        Bytes blob;
        std::vector<uint64_t> off(props.size() + 1, 0);
        for (size_t i = 0; i < props.size(); i++) {
            Bytes d = props[i].Der();
            blob.insert(blob.end(), d.begin(), d.end());
            off[i + 1] = blob.size();
        }
        Bytes dig(props.size() * 32);
        if (sbv_sha256_batch(eng_, props.size(), blob.data(), off.data(), dig.data()) != SBV_OK)
            throw EngineFault(std::string("sbv_sha256_batch: ") + sbv_last_error(eng_));
        for (size_t i = 0; i < props.size(); i++) out[i] = hex(Bytes(dig.begin() + 32 * i, dig.begin() + 32 * i + 32));
        return out;
    }

    // ---- api.Verifier ----
    std::pair<Bytes, Error> VerifyConsenterSig(const Signature &sig, const Proposal &prop) override {
        SigItem it;
        if (Error e = prepare_consenter(sig, prop.DigestRaw(), it)) return {Bytes(), e};
        if (!agg_->submit(std::move(it))) return {Bytes(), Errorf("invalid signature from " + std::to_string(sig.ID))};
        return {AuxiliaryData(sig.Msg), std::nullopt};
    }
    std::vector<std::pair<Bytes, Error>> VerifyConsenterSigBatch(const std::vector<Signature> &sigs, const Proposal &prop) override {
        const Bytes dig = prop.DigestRaw();
        std::vector<std::pair<Bytes, Error>> out(sigs.size());
        std::vector<SigItem> items;
        std::vector<size_t> where;
        for (size_t i = 0; i < sigs.size(); i++) {
            SigItem it;
            if (Error e = prepare_consenter(sigs[i], dig, it)) { out[i] = {Bytes(), e}; continue; }
            items.push_back(std::move(it)); where.push_back(i);
        }
        if (!items.empty()) {
            auto ok = engine_batch(items);
            for (size_t k = 0; k < where.size(); k++) {
                size_t i = where[k];
                if (ok[k]) out[i] = {AuxiliaryData(sigs[i].Msg), std::nullopt};
                else out[i] = {Bytes(), Errorf("invalid signature from " + std::to_string(sigs[i].ID))};
            }
        }
        return out;
    }
    Error VerifySignature(const Signature &sig) override {
        SigItem it;
        if (Error e = prepare_plain(sig, it)) return e;
        if (!agg_->submit(std::move(it))) return Errorf("invalid signature from " + std::to_string(sig.ID));
        return std::nullopt;
    }
    std::pair<RequestInfo, Error> VerifyRequest(const Bytes &val) override {
        SigItem it; RequestInfo info;
        if (Error e = prepare_request(val, it, info)) return {RequestInfo(), e};
        if (!agg_->submit(std::move(it))) return {RequestInfo(), Errorf("bad request signature")};
        return {info, std::nullopt};
    }
    std::vector<std::pair<RequestInfo, Error>> VerifyRequestBatch(const std::vector<Bytes> &reqs) override {
        std::vector<std::pair<RequestInfo, Error>> out(reqs.size());
        std::vector<SigItem> items; std::vector<size_t> where; std::vector<RequestInfo> infos;
        for (size_t i = 0; i < reqs.size(); i++) {
            SigItem it; RequestInfo info;
            if (Error e = prepare_request(reqs[i], it, info)) { out[i] = {RequestInfo(), e}; continue; }
            items.push_back(std::move(it)); where.push_back(i); infos.push_back(info);
        }
        if (!items.empty()) {
            auto ok = engine_batch(items);
            for (size_t k = 0; k < where.size(); k++)
                out[where[k]] = ok[k] ? std::make_pair(infos[k], Error()) : std::make_pair(RequestInfo(), Errorf("bad request signature"));
        }
        return out;
    }
    // view.go:555 — verifies every request of the batch in ONE engine call; any failure rejects.
    std::pair<std::vector<RequestInfo>, Error> VerifyProposal(const Proposal &proposal) override {
        std::vector<Bytes> reqs;
        if (!split_requests(proposal.Payload, reqs)) return {{}, Errorf("malformed proposal payload")};
        {
            std::lock_guard<std::mutex> lk(mu_);
            if ((uint64_t)proposal.VerificationSequence != verSeq_) return {{}, Errorf("verification sequence mismatch")};
        }
        auto res = VerifyRequestBatch(reqs);
        std::vector<RequestInfo> infos;
        for (auto &p : res) { if (p.second) return {{}, p.second}; infos.push_back(p.first); }
        return {infos, std::nullopt};
    }
    uint64_t VerificationSequence() override { std::lock_guard<std::mutex> lk(mu_); return verSeq_; }
    std::vector<RequestInfo> RequestsFromProposal(const Proposal &proposal) override {
        std::vector<Bytes> reqs; std::vector<RequestInfo> infos;
        if (!split_requests(proposal.Payload, reqs)) return infos;
        for (auto &r : reqs) { ParsedRequest pr; if (parse_request(r, pr)) infos.push_back({pr.client, pr.id}); }
        return infos;
    }
    Bytes AuxiliaryData(const Bytes &msg) override { return msg.size() >= 32 ? Bytes(msg.begin() + 32, msg.end()) : Bytes(); }

  private:
    Error prepare_plain(const Signature &sig, SigItem &it) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            auto k = consenters_.find(sig.ID);
            if (k == consenters_.end()) return Errorf("unknown consenter " + std::to_string(sig.ID));
            it.slot = k->second;
        }
        if (!parse_der_sig(sig.Value, it.r, it.s)) return Errorf("malformed signature from " + std::to_string(sig.ID));
        it.message = sig.Msg;
        return std::nullopt;
    }
    Error prepare_consenter(const Signature &sig, const Bytes &digest, SigItem &it) {
        if (sig.Msg.size() < 32 || memcmp(sig.Msg.data(), digest.data(), 32) != 0) return Errorf("signature does not bind the proposal");
        return prepare_plain(sig, it);
    }
    Error prepare_request(const Bytes &val, SigItem &it, RequestInfo &info) {
        ParsedRequest pr;
        if (!parse_request(val, pr)) return Errorf("malformed request");
        {
            std::lock_guard<std::mutex> lk(mu_);
            auto k = clients_.find(pr.client);
            if (k == clients_.end()) return Errorf("unknown client " + pr.client);
            it.slot = k->second;
        }
        if (!parse_der_sig(pr.sig, it.r, it.s)) return Errorf("malformed request signature");
        it.message = pr.signedBytes;
        info = {pr.client, pr.id};
        return std::nullopt;
    }
    uint32_t slot_of(const uint8_t xy[64]) {  // mu_ held
        std::array<uint8_t, 64> k;
        memcpy(k.data(), xy, 64);
        auto f = slots_.find(k);
        if (f != slots_.end()) return f->second;
        uint32_t s = (uint32_t)registry_.size();
        registry_.push_back(k);
        slots_[k] = s;
        dirty_ = true;
        return s;
    }
    void sync_registry() {
        std::lock_guard<std::mutex> lk(mu_);
        if (!dirty_) return;
        const size_t n = registry_.size();
        std::vector<uint64_t> ids(n);
        std::vector<uint8_t> curve(n, SBV_P256), xy(n * 96, 0);
        for (size_t i = 0; i < n; i++) {
            ids[i] = i;
            memcpy(&xy[96 * i + 16], registry_[i].data(), 32);
            memcpy(&xy[96 * i + 48 + 16], registry_[i].data() + 32, 32);
        }
        if (sbv_set_keys(eng_, verSeq_, n, ids.data(), curve.data(), xy.data()) != SBV_OK)
            throw EngineFault(std::string("sbv_set_keys: ") + sbv_last_error(eng_));
        dirty_ = false;
    }
    sbv_engine *eng_ = nullptr;
    std::unique_ptr<Aggregator> agg_;
    std::mutex mu_;
    uint64_t verSeq_ = 0;
    bool dirty_ = false;
    std::vector<std::array<uint8_t, 64>> registry_;
    std::map<std::array<uint8_t, 64>, uint32_t> slots_;
    std::map<uint64_t, uint32_t> consenters_;
    std::map<std::string, uint32_t> clients_;
};

}  // namespace sbft
