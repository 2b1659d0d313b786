// CPU oracle #2 (C++17, multithreaded) for classeq's placement hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.  It shares no source with the
// product (classeq2_b200/): own murmur3, own model containers, own descent - written in the
// reference's SET formulation (HashSet unions / differences), not the counter formulation the
// GPU kernel uses, so that agreement between the two means something.
//
// Function by function it restates the reference's Rust (paths relative to its checkout):
//   murmur3_x64_128            crate mur3 0.1.0 (Cargo.lock:2405-2408; not vendored) = published
//                              MurmurHash3_x64_128; sole call site core/src/domain/dtos/kmers_map.rs:157-159
//   build_kmers                kmers_map.rs:375-398 (+ :405-424 windows, :431-443 reverse complement)
//   overlapping (M)            kmers_map.rs:273-311 (+ :55-70): entry (bucket, hash) survives iff
//                              bucket in {h1(prefix_m(w))} and hash in {h1(w)} over the query windows w
//   root restriction (M_r)     kmers_map.rs:211-229, :318-344; place_sequence.rs:156-166
//   K(node)                    kmers_map.rs:189-203 (HashSet<u64> of hashes, flattened over buckets)
//   place                      core/src/use_cases/place_sequences/place_sequence.rs:42-602
//   descend                    core/src/use_cases/place_sequences/update_introspection_node.rs:13-91
//
// One deliberate difference in COST, none in RESULT: the reference clones the whole index per
// query and scans every bucket (place_sequence.rs:77-80, kmers_map.rs:55-70); this oracle looks
// each query hash up directly.  It is therefore a faster - i.e. conservative - CPU baseline.
//
// PARITY STATUS: "parity unpinned" for placement decisions (the reference is Rust, cannot be built
// here, and its placement goldens depend on a missing Git-LFS model); pinned for the hash, the
// windowing and the both-strands count (tests/test_oracle_kats.py).  This file and
// oracle/classeq_oracle.py are two independent restatements that tests hold equal.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

// ---- MurmurHash3_x64_128 (A. Appleby, public domain algorithm), byte-wise loads -----------------
inline uint64_t rotl(uint64_t x, unsigned r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t fmix(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}
inline uint64_t getle(const unsigned char *p, int n) {
    uint64_t v = 0;
    for (int i = n - 1; i >= 0; --i) v = (v << 8) | p[i];
    return v;
}
void murmur128(const unsigned char *d, uint64_t n, uint64_t seed, uint64_t out[2]) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    uint64_t i = 0;
    for (; i + 16 <= n; i += 16) {
        uint64_t k1 = getle(d + i, 8), k2 = getle(d + i + 8, 8);
        k1 *= c1; k1 = rotl(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const int t = (int)(n - i);
    if (t > 8) { uint64_t k2 = getle(d + i + 8, t - 8); k2 *= c2; k2 = rotl(k2, 33); k2 *= c1; h2 ^= k2; }
    if (t > 0) { uint64_t k1 = getle(d + i, t > 8 ? 8 : t); k1 *= c1; k1 = rotl(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= n; h2 ^= n;
    h1 += h2; h2 += h1;
    h1 = fmix(h1); h2 = fmix(h2);
    h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}
inline uint64_t h1_of(const unsigned char *d, uint64_t n) { uint64_t o[2]; murmur128(d, n, 0, o); return o[0]; }

// ---- model ------------------------------------------------------------------------------------------
struct Clade {
    uint64_t id = 0;
    int kind = 1;                  // 0 ROOT, 1 NODE, 2 LEAF  (clade.rs:5-16)
    std::vector<int> children;     // indices into Model::clades, Clade.children order
};
struct Entry { uint64_t bucket; uint32_t set; };
struct Model {
    uint32_t k = 0, m = 0;
    bool root_children_none = false;
    std::vector<Clade> clades;     // clades[0] = tree.root
    std::vector<std::vector<uint64_t>> sets;  // sorted node ids
    std::unordered_map<uint64_t, std::vector<Entry>> by_hash;  // hash -> (bucket key, set) of every bucket holding it
};
inline bool set_has(const std::vector<uint64_t> &s, uint64_t id) { return std::binary_search(s.begin(), s.end(), id); }

struct Hit { uint64_t bucket, hash; uint32_t set; };

enum Status : uint8_t {  // numeric values follow include/classeq_b200.h's cls_status for the tests' convenience
    ERR_TOO_SHORT = 0, UNCL_NO_MATCH = 1, UNCL_NO_ROOT = 2, UNCL_COVERAGE = 3, UNCL_NO_INTROSPECTION = 4,
    MAX_RESOLUTION = 5, IDENTITY_FOUND = 6, INCONCLUSIVE = 7, ERR_MAX_ITERATIONS = 8, ERR_ROOT_NO_CHILDREN = 9,
    ERR_INVALID_BASE = 10
};
struct Outcome {
    uint8_t status = 0;
    uint64_t node = 0;
    int32_t one = 0, rest = 0;
    uint32_t n_query = 0, n_matched = 0, n_root = 0, iterations = 0;
};

// kmers_map.rs:375-398: all forward windows, then all windows of the reverse complement
bool build_kmers(const Model &md, const unsigned char *seq, uint64_t len, std::vector<uint64_t> &hashes,
                 std::vector<uint64_t> &prefix_keys, bool &invalid) {
    invalid = false;
    hashes.clear(); prefix_keys.clear();
    if (len < md.k) return false;  // kmers_map.rs:383-385
    std::string f(len, 'A'), r(len, 'A');
    for (uint64_t i = 0; i < len; ++i) {
        unsigned char c = seq[i];
        if (c >= 'a' && c <= 'z') c = (unsigned char)(c - 32);  // :410 to_uppercase
        f[i] = (char)c;
        char rc;
        switch (c) {  // :431-443
            case 'A': rc = 'T'; break; case 'T': rc = 'A'; break;
            case 'C': rc = 'G'; break; case 'G': rc = 'C'; break;
            default: invalid = true; return false;  // reference: panic!("Invalid character in sequence")
        }
        r[len - 1 - i] = rc;
    }
    const uint64_t W = len - md.k + 1;
    const uint64_t mm = std::min<uint64_t>(md.m, md.k);  // chars().take(m) of a k-char k-mer
    for (int strand = 0; strand < 2; ++strand) {
        const unsigned char *s = (const unsigned char *)(strand ? r.data() : f.data());
        for (uint64_t i = 0; i < W; ++i) {
            hashes.push_back(h1_of(s + i, md.k));       // :157-159
            prefix_keys.push_back(h1_of(s + i, mm));    // :10-13 (h1("") = 0 when m = 0)
        }
    }
    return true;
}

Outcome place(const Model &md, const unsigned char *seq, uint64_t len, int32_t max_iter, double cov, bool ri) {
    Outcome o;
    std::vector<uint64_t> hashes, pkeys;
    bool invalid;
    if (!build_kmers(md, seq, len, hashes, pkeys, invalid) || hashes.size() < 2) {  // place_sequence.rs:98-102
        o.status = invalid ? ERR_INVALID_BASE : ERR_TOO_SHORT;
        return o;
    }
    o.n_query = (uint32_t)hashes.size();
    // kmers_map.rs:273-311
    std::unordered_set<uint64_t> minimizers(pkeys.begin(), pkeys.end());
    std::unordered_set<uint64_t> qh(hashes.begin(), hashes.end());
    std::vector<Hit> M;
    for (uint64_t h : qh) {
        auto it = md.by_hash.find(h);
        if (it == md.by_hash.end()) continue;
        for (const Entry &e : it->second)
            if (minimizers.count(e.bucket)) M.push_back(Hit{e.bucket, h, e.set});
    }
    o.n_matched = (uint32_t)M.size();  // sum over buckets of bucket sizes (place_sequence.rs:120-125)
    if (M.empty()) { o.status = UNCL_NO_MATCH; return o; }
    // place_sequence.rs:156-166
    const uint64_t root_id = md.clades[0].id;
    std::vector<Hit> Mr;
    for (const Hit &h : M) if (set_has(md.sets[h.set], root_id)) Mr.push_back(h);
    if (Mr.empty()) { o.status = UNCL_NO_ROOT; return o; }
    o.n_root = (uint32_t)Mr.size();
    if (md.root_children_none) { o.status = ERR_ROOT_NO_CHILDREN; return o; }  // :199-206
    // :231-254   f64::round = half away from zero = std::round
    const double expected = std::round((double)M.size() * cov);
    const double clamped = expected < 0 ? 0 : expected;  // `as usize` saturates
    if ((double)Mr.size() < clamped) { o.status = UNCL_COVERAGE; return o; }

    int parent = 0;
    std::vector<int> children = md.clades[0].children;
    int64_t iteration = 0;
    for (;;) {
        iteration++;
        o.iterations = (uint32_t)iteration;
        if (iteration > (int64_t)max_iter) { o.status = ERR_MAX_ITERATIONS; return o; }  // :295-301
        // PHASE 1 (:311-428)
        std::vector<std::pair<std::unordered_set<uint64_t>, int>> ck;
        for (int c : children) {
            if (md.clades[c].kind == 2) continue;  // is_leaf() is by kind
            std::unordered_set<uint64_t> K;
            for (const Hit &h : Mr) if (set_has(md.sets[h.set], md.clades[c].id)) K.insert(h.hash);
            if (!K.empty()) ck.emplace_back(std::move(K), c);
        }
        struct Prop { int clade; int32_t one, rest; };
        std::vector<Prop> props;
        for (size_t a = 0; a < ck.size(); ++a) {
            std::unordered_set<uint64_t> rest;
            size_t n_rest_sets = 0;
            for (size_t b = 0; b < ck.size(); ++b) {
                if (md.clades[ck[b].second].id == md.clades[ck[a].second].id) continue;
                n_rest_sets++;
                rest.insert(ck[b].first.begin(), ck[b].first.end());
            }
            int32_t one, rst;
            if (n_rest_sets == 0) { one = (int32_t)ck[a].first.size(); rst = 0; }
            else if (ri) {
                size_t inter = 0;
                for (uint64_t x : ck[a].first) inter += rest.count(x);
                one = (int32_t)(ck[a].first.size() - inter);
                rst = (int32_t)(rest.size() - inter);
            } else { one = (int32_t)ck[a].first.size(); rst = (int32_t)rest.size(); }
            if (one > rst) props.push_back(Prop{ck[a].second, one, rst});
        }
        // PHASE 2 (:436-600)
        if (props.empty()) {
            if (iteration == 1) { o.status = UNCL_NO_INTROSPECTION; return o; }
            o.status = MAX_RESOLUTION; o.node = md.clades[parent].id; return o;
        }
        Prop win = props[0];
        if (props.size() > 1) {
            int32_t best = INT32_MIN; size_t nbest = 0;
            for (const Prop &p : props) {
                int32_t d = p.one - p.rest;
                if (d > best) { best = d; nbest = 1; win = p; } else if (d == best) nbest++;
            }
            if (nbest != 1) { o.status = INCONCLUSIVE; o.node = md.clades[parent].id; return o; }
        }
        // update_introspection_node.rs:13-91
        std::vector<int> nl;
        for (int c : md.clades[win.clade].children) if (md.clades[c].kind != 2) nl.push_back(c);
        if (nl.empty()) {
            o.status = IDENTITY_FOUND; o.node = md.clades[win.clade].id; o.one = win.one; o.rest = win.rest;
            return o;
        }
        parent = win.clade;
        children = nl;
    }
}

}  // namespace

extern "C" {

// Same field order as cls_model_view (include/classeq_b200.h) so that the tests can hand the very
// same ctypes structure to both sides.
struct orc_model_view {
    uint32_t k_size, m_size, flags, reserved;
    uint64_t n_nodes;
    const uint64_t *node_id; const uint8_t *node_kind; const uint64_t *child_off; const uint64_t *child_idx;
    uint64_t n_entries;
    const uint64_t *entry_bucket; const uint64_t *entry_hash; const uint64_t *entry_set;
    uint64_t n_sets;
    const uint64_t *set_off; const uint64_t *set_node_ids;
};

void *orc_model_create(const orc_model_view *v) {
    auto *md = new Model();
    md->k = v->k_size; md->m = v->m_size; md->root_children_none = (v->flags & 1u) != 0;
    md->clades.resize(v->n_nodes);
    for (uint64_t i = 0; i < v->n_nodes; ++i) {
        md->clades[i].id = v->node_id[i];
        md->clades[i].kind = v->node_kind[i];
        for (uint64_t j = v->child_off[i]; j < v->child_off[i + 1]; ++j) md->clades[i].children.push_back((int)v->child_idx[j]);
    }
    md->sets.resize(v->n_sets);
    for (uint64_t s = 0; s < v->n_sets; ++s) {
        md->sets[s].assign(v->set_node_ids + v->set_off[s], v->set_node_ids + v->set_off[s + 1]);
        std::sort(md->sets[s].begin(), md->sets[s].end());
    }
    md->by_hash.reserve(v->n_entries * 2);
    for (uint64_t e = 0; e < v->n_entries; ++e)
        md->by_hash[v->entry_hash[e]].push_back(Entry{v->entry_bucket[e], (uint32_t)v->entry_set[e]});
    return md;
}
void orc_model_destroy(void *m) { delete (Model *)m; }

// Places queries [0, n) with `n_threads` workers, one query per task (the reference's par_bridge
// over queries, place_sequences/mod.rs:123-126).  Knobs: place_sequence.rs:64-75.
void orc_place_batch(const void *model, const uint8_t *bases, const uint64_t *offsets, uint64_t n,
                     int32_t max_iterations, double min_match_coverage, uint32_t remove_intersection, int n_threads,
                     uint8_t *status, uint64_t *node_id, int32_t *one, int32_t *rest, uint32_t *n_query_kmers,
                     uint32_t *n_matched, uint32_t *n_root_matched, uint32_t *iterations) {
    const Model &md = *(const Model *)model;
    double cov = min_match_coverage;
    if (cov > 1.0) cov = 1.0; else if (cov < 0.0) cov = 0.0;  // NaN passes through, as in the reference
    std::atomic<uint64_t> next{0};
    auto worker = [&] {
        for (;;) {
            const uint64_t a = next.fetch_add(64);
            if (a >= n) break;
            const uint64_t b = std::min(n, a + 64);
            for (uint64_t i = a; i < b; ++i) {
                Outcome o = place(md, bases + offsets[i], offsets[i + 1] - offsets[i], max_iterations, cov, remove_intersection != 0);
                status[i] = o.status; node_id[i] = o.node; one[i] = o.one; rest[i] = o.rest;
                n_query_kmers[i] = o.n_query; n_matched[i] = o.n_matched; n_root_matched[i] = o.n_root; iterations[i] = o.iterations;
            }
        }
    };
    if (n_threads <= 1) { worker(); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(worker);
    for (auto &t : th) t.join();
}

// ---- the k-mer map of a tree from its tips' sequences ------------------------------------------------------
// Restates build_database/mod.rs:62 and :140-168 (every k-mer of a tip's sequence, both strands, is inserted with the
// ids on the root -> tip path: `get_leaves_with_paths`, `insert_or_append_kmer_hash`), for callers that pair every tip
// with its own sequence.  Own implementation (sort + group), shares nothing with the product's builders.
// Output as flat arrays (orc_built_*): entries sorted by (hash, bucket); distinct node sets, ids ascending.
struct Built {
    std::vector<uint64_t> entry_bucket, entry_hash, entry_set, set_off, set_node_ids;
};

void *orc_model_build(uint32_t k_size, uint32_t m_size, uint64_t n_nodes, const uint64_t *node_id, const uint64_t *child_off,
                      const uint64_t *child_idx, uint64_t n_tips, const uint64_t *tip_node, const uint8_t *bases,
                      const uint64_t *offsets, int n_threads) {
    struct Occ { uint64_t hash, bucket; uint32_t tip; };
    Model probe; probe.k = k_size; probe.m = m_size;
    std::vector<std::vector<Occ>> per_thread((size_t)std::max(1, n_threads));
    std::atomic<uint64_t> next{0};
    std::atomic<bool> bad{false};
    auto worker = [&](int w) {
        std::vector<uint64_t> h, pk; bool inv;
        for (;;) {
            const uint64_t t = next.fetch_add(1);
            if (t >= n_tips) break;
            const unsigned char *sq = bases + offsets[t];
            const uint64_t len = offsets[t + 1] - offsets[t];
            if (!build_kmers(probe, sq, len, h, pk, inv)) { if (inv) bad = true; continue; }
            // build_kmers lists the forward windows, then the windows of the reverse complement; the bucket key of a
            // window is the hash of its first m characters (kmers_map.rs:10-13), i.e. pk in the same order
            for (size_t i = 0; i < h.size(); ++i) per_thread[(size_t)w].push_back(Occ{h[i], pk[i], (uint32_t)t});
        }
    };
    {
        std::vector<std::thread> th;
        for (int w = 0; w < std::max(1, n_threads); ++w) th.emplace_back(worker, w);
        for (auto &t : th) t.join();
    }
    if (bad) return nullptr;
    std::vector<Occ> occ;
    { size_t n = 0; for (auto &v : per_thread) n += v.size(); occ.reserve(n); }
    for (auto &v : per_thread) { occ.insert(occ.end(), v.begin(), v.end()); std::vector<Occ>().swap(v); }
    std::sort(occ.begin(), occ.end(), [](const Occ &a, const Occ &b) {
        return a.hash != b.hash ? a.hash < b.hash : (a.bucket != b.bucket ? a.bucket < b.bucket : a.tip < b.tip);
    });
    // parent of every node, for the root -> tip paths
    std::vector<int64_t> parent(n_nodes, -1);
    for (uint64_t v = 0; v < n_nodes; ++v)
        for (uint64_t j = child_off[v]; j < child_off[v + 1]; ++j) parent[child_idx[j]] = (int64_t)v;
    auto *out = new Built();
    struct VecHash { size_t operator()(const std::vector<uint32_t> &v) const { size_t x = 1469598103934665603ull; for (uint32_t e : v) { x ^= e; x *= 1099511628211ull; } return x; } };
    std::unordered_map<std::vector<uint32_t>, uint64_t, VecHash> set_of_tips;
    out->set_off.push_back(0);
    std::vector<uint32_t> tips;
    for (size_t i = 0; i < occ.size();) {
        size_t j = i;
        tips.clear();
        while (j < occ.size() && occ[j].hash == occ[i].hash && occ[j].bucket == occ[i].bucket) {
            if (tips.empty() || tips.back() != occ[j].tip) tips.push_back(occ[j].tip);
            ++j;
        }
        auto it = set_of_tips.find(tips);
        uint64_t sidx;
        if (it == set_of_tips.end()) {
            sidx = out->set_off.size() - 1;
            set_of_tips.emplace(tips, sidx);
            std::vector<uint64_t> ids;
            for (uint32_t t : tips)
                for (int64_t v = (int64_t)tip_node[t]; v >= 0; v = parent[(size_t)v]) ids.push_back(node_id[(size_t)v]);
            std::sort(ids.begin(), ids.end());
            ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
            out->set_node_ids.insert(out->set_node_ids.end(), ids.begin(), ids.end());
            out->set_off.push_back(out->set_node_ids.size());
        } else {
            sidx = it->second;
        }
        out->entry_bucket.push_back(occ[i].bucket); out->entry_hash.push_back(occ[i].hash); out->entry_set.push_back(sidx);
        i = j;
    }
    return out;
}
void orc_built_sizes(const void *b, uint64_t *n_entries, uint64_t *n_sets, uint64_t *n_ids) {
    const Built &x = *(const Built *)b;
    *n_entries = x.entry_hash.size(); *n_sets = x.set_off.size() - 1; *n_ids = x.set_node_ids.size();
}
void orc_built_copy(const void *b, uint64_t *entry_bucket, uint64_t *entry_hash, uint64_t *entry_set, uint64_t *set_off, uint64_t *set_node_ids) {
    const Built &x = *(const Built *)b;
    std::copy(x.entry_bucket.begin(), x.entry_bucket.end(), entry_bucket);
    std::copy(x.entry_hash.begin(), x.entry_hash.end(), entry_hash);
    std::copy(x.entry_set.begin(), x.entry_set.end(), entry_set);
    std::copy(x.set_off.begin(), x.set_off.end(), set_off);
    std::copy(x.set_node_ids.begin(), x.set_node_ids.end(), set_node_ids);
}
void orc_built_destroy(void *b) { delete (Built *)b; }

// All window hashes of one query, reference order (forward, then reverse complement).
uint64_t orc_kmer_hashes(const uint8_t *bases, uint64_t len, uint32_t k, uint64_t *out, uint64_t cap) {
    Model md; md.k = k; md.m = 0;
    std::vector<uint64_t> h, p; bool inv;
    if (!build_kmers(md, bases, len, h, p, inv)) return 0;
    for (uint64_t i = 0; i < h.size() && i < cap; ++i) out[i] = h[i];
    return h.size();
}

void orc_murmur3_x64_128(const uint8_t *data, uint64_t len, uint64_t seed, uint64_t out[2]) { murmur128(data, len, seed, out); }

}  // extern "C"
