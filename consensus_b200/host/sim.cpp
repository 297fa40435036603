// sim.cpp — n=4 (f=1, Q=3) in-process normal-path simulator for the "consensus tx/s" metric
// (BASELINE configs[0]: examples/naive_chain, 1K tx).  BENCH INFRASTRUCTURE.
//
// Restates the normal case of View only (pre-prepare -> Q-1 prepares -> Q-1 verified commits ->
// decide; /root/reference/internal/bft/view.go:282-299, 351-551, 851-894) with the verifier calls
// every node makes per decision, and no WAL / view change / networking:
//   leader ingress       controller.go:233-246   VerifyRequest per client request (B per batch)
//   3 followers          view.go:553-604         VerifyProposal (B requests) + verifyPrevCommitSignatures (Q sigs)
//   4 nodes              view.go:519-551         processCommits over the N-1 foreign commit votes
// All four nodes share ONE verifier instance (one GPU); a deployment has one per node, so the GPU
// figure is a lower bound.  The nodes of a phase run on their own threads, as the processes of a deployment would.  Signatures are produced up front (signing is api.Signer's job).
#include <atomic>
#include <chrono>
#include <thread>
#include <cstdio>
#include <cstdlib>

#include "test_signer.hpp"

struct Decision { Proposal prop; std::vector<Bytes> reqs; std::vector<Vote> commits; std::vector<Vote> prepares; std::vector<ProtoSignature> quorumSigs; };

static double run(IVerifier &v, const std::vector<Decision> &ds, int N, int Q, size_t &txs, bool &allDecided) {
    auto t0 = std::chrono::steady_clock::now();
    txs = 0; allDecided = true;
    Pool pool;
    const Decision *prev = nullptr;
    for (const Decision &d : ds) {
        // leader ingress: every request is verified, then pooled (HandleRequest; batched at the pool boundary)
        auto in = v.VerifyRequestBatch(d.reqs);
        for (auto &p : in) if (p.second) allDecided = false;
        // the nodes of a deployment run concurrently (one process each): followers verify the pre-prepare in
        // parallel, then every node collects its commit votes in parallel; a phase ends when its slowest node does
        std::atomic<bool> good{true};
        {
            std::vector<std::thread> th;
            for (uint64_t node = 2; node <= (uint64_t)N; node++)
                th.emplace_back([&] {
                    auto vp = v.VerifyProposal(d.prop);
                    if (vp.second) good = false;
                    if (prev) { auto pc = verifyPrevCommitSignatures(v, prev->quorumSigs, prev->prop, 1); if (pc.second) good = false; }
                });
            for (auto &t : th) t.join();
        }
        {
            std::vector<std::thread> th;
            for (uint64_t node = 1; node <= (uint64_t)N; node++)
                th.emplace_back([&, node] {
                    auto ids = processPrepares(d.prop, Q, node, d.prepares);
                    auto sigs = processCommits(v, d.prop, Q, node, d.commits);
                    if ((int)ids.size() != Q - 1 || (int)sigs.size() != Q - 1) good = false;
                });
            for (auto &t : th) t.join();
        }
        if (!good) allDecided = false;
        txs += d.reqs.size();
        prev = &d;
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char **argv) {
    const int N = 4, TX = argc > 1 ? atoi(argv[1]) : 1000, B = argc > 2 ? atoi(argv[2]) : 100;
    const bool gpu = argc > 3 ? atoi(argv[3]) != 0 : true;
    int Q, F; computeQuorum(N, Q, F);
    std::map<uint64_t, TestKey> keys;
    for (uint64_t id = 1; id <= (uint64_t)N; id++) keys[id] = makeKey();
    TestKey ck = makeKey();
    // pre-generate the run: requests, proposals, prepares, signed commits
    std::vector<Decision> ds;
    int txid = 0;
    for (int seq = 0; txid < TX; seq++) {
        Decision d;
        Bytes payload;
        for (int k = 0; k < B && txid < TX; k++, txid++) {
            Bytes body(256, (uint8_t)txid);
            Bytes sig = signDer(ck, signed_part("alice", std::to_string(txid), body));
            Bytes req = frame_request(sig, "alice", std::to_string(txid), body);
            for (int s = 3; s >= 0; s--) payload.push_back((uint8_t)(req.size() >> (8 * s)));
            payload.insert(payload.end(), req.begin(), req.end());
            d.reqs.push_back(req);
        }
        d.prop = Proposal{payload, {(uint8_t)seq}, ViewMetadata{1, (uint64_t)seq, (uint64_t)seq}.Marshal(), 1};
        Bytes aux = PreparesFrom{{2, 3}}.Marshal();
        std::string dg = d.prop.Digest();
        for (uint64_t id = 1; id <= (uint64_t)N; id++) {
            Vote p; p.sender = id; p.prepare = PrepareMsg{1, (uint64_t)seq, dg, false};
            d.prepares.push_back(p);
            Signature s = signProposal(id, keys[id], d.prop, aux);
            Vote c; c.sender = id; CommitMsg cm; cm.View = 1; cm.Seq = seq; cm.Digest = dg; cm.Sig = ProtoSignature{id, s.Value, s.Msg}; c.commit = cm;
            d.commits.push_back(c);
            if ((int)d.quorumSigs.size() < Q) d.quorumSigs.push_back(ProtoSignature{id, s.Value, s.Msg});
        }
        ds.push_back(std::move(d));
    }
    size_t txs; bool ok;
    AcceptAllVerifier acc;
    double ta = run(acc, ds, N, Q, txs, ok);
    printf("{\"n\": %d, \"f\": %d, \"quorum\": %d, \"txs\": %zu, \"batch\": %d, \"decisions\": %zu", N, F, Q, txs, B, ds.size());
    printf(", \"accept_all_tx_per_s\": %.1f", txs / ta);
    CpuVerifier cpu;
    for (auto &kv : keys) cpu.consenters[kv.first] = kv.second.k;
    cpu.clients["alice"] = ck.k;
    double tc = run(cpu, ds, N, Q, txs, ok);
    printf(", \"cpu_per_call_tx_per_s\": %.1f, \"cpu_all_decided\": %s", txs / tc, ok ? "true" : "false");
    if (gpu) {
        GpuVerifier g({0});
        g.SetVerificationSequence(1);
        for (auto &kv : keys) g.SetConsenterKey(kv.first, kv.second.xy);
        g.SetClientKey("alice", ck.xy);
        run(g, ds, N, Q, txs, ok);  // warm-up (allocations, table residency)
        double tg = run(g, ds, N, Q, txs, ok);
        printf(", \"gpu_tx_per_s\": %.1f, \"gpu_all_decided\": %s, \"gpu_engine_calls\": %llu", txs / tg, ok ? "true" : "false",
               (unsigned long long)sbv_kernel_launches(g.engine()));
    }
    printf(", \"verifications_per_decision\": %d}\n", 4 * B + 3 * Q + N * (N - 1));
    return 0;
}
