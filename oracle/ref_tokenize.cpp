// TEST INFRASTRUCTURE (oracle) — not product code.
//
// Faithful C++ restatement of the reference's CPU hot path (togatoga/kanpyo @ f1931e2c).  It keeps
// the reference's *shape* on purpose (per-node heap surface string + morph copy, vector<vector<>>
// end-position buckets, per-call result vectors, Option-like dp) because it doubles as the
// "restated reference CPU path" timed by bench.py's cpu_baseline / --impl reference legs.
//
//   DoubleArray::search                    kanpyo-dict/src/trie/da.rs:133-153
//   DoubleArray::search_common_prefix_of   kanpyo-dict/src/trie/da.rs:155-182
//   IndexTable::search_common_prefix_of    kanpyo-dict/src/index.rs:40-53
//   ConnectionTable::get                   kanpyo-dict/src/connection.rs:12-14
//   CharCategoryDef::char_category         kanpyo-dict/src/char_category_def.rs:33-38
//   Morphs index (id-1)                    kanpyo-dict/src/morph.rs:46-52
//   Lattice::{new,build,process_*,add_*}   src/lattice.rs:13-114,156-201
//   Lattice::viterbi                       src/lattice.rs:116-154
//   Node accessors                         src/lattice/node.rs:27-52
//   Tokenizer::tokenize                    src/tokenizer.rs:16-45
//
// Parity pin: tests/test_oracle_pins.py checks this file against every known-answer test the
// reference holds for the path (da.rs:253-351, index.rs:92-150, connection.rs:58-72,
// matrix_def.rs:70-85, src/tests.rs fixture) and the README end-to-end outputs.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <optional>
#include <string>
#include <string_view>
#include <thread>
#include <utility>
#include <vector>

namespace {

constexpr uint8_t TERMINATOR = 0;  // da.rs:8
constexpr size_t ROOT_ID = 1;      // da.rs:9
constexpr int64_t BOS_EOS_ID = 0;  // node.rs:3

struct DaNode {  // da.rs:14-17
    int32_t base;
    int32_t check;
};

struct Morph {  // morph.rs:7-11
    int16_t left_id, right_id, cost;
};

struct Counters {
    uint64_t bytes = 0, chars = 0, probes = 0, probes_ok = 0, nodes = 0, pairs = 0, tokens = 0;
};

struct Dict {  // dict.rs:21-30 (hot-path members only)
    std::vector<DaNode> da;                                     // index.rs:11
    std::map<int64_t, size_t> dup;                              // index.rs:12 (BTreeMap)
    std::vector<Morph> morphs;                                  // morph.rs:24
    size_t conn_row = 0, conn_col = 0;                          // connection.rs:5-9
    std::vector<int16_t> conn;
    std::vector<uint8_t> char_category;                         // char_category_def.rs:15-20
    std::vector<bool> invoke_list, group_list;
    std::map<uint8_t, std::pair<int64_t, size_t>> unk_map;      // unk_dict.rs:15
    std::vector<Morph> unk_morphs;                              // unk_dict.rs:13

    // Vec::get with a usize index: a negative i32 cast to usize is huge => None.
    const DaNode* da_get(int32_t i) const {
        if (i < 0 || (size_t)i >= da.size()) return nullptr;
        return &da[(size_t)i];
    }
    int16_t conn_get(size_t row, size_t col) const { return conn.at(conn_row * col + row); }  // connection.rs:12-14
    uint8_t category(uint32_t ch) const {  // char_category_def.rs:33-38
        return ch < char_category.size() ? char_category[ch] : char_category.at(0);
    }
};

// da.rs:133-153
std::optional<int64_t> da_search(const Dict& d, std::string_view keyword) {
    int32_t p = (int32_t)ROOT_ID;
    for (unsigned char ch : keyword) {
        const DaNode* n = d.da_get(p);
        if (!n) return std::nullopt;
        int32_t q = n->base + (int32_t)ch;
        const DaNode* m = d.da_get(q);
        if (!m) return std::nullopt;
        if (m->check != p) return std::nullopt;
        p = q;
    }
    const DaNode* n = d.da_get(p);
    if (!n) return std::nullopt;
    int32_t q = n->base + TERMINATOR + TERMINATOR;
    const DaNode* m = d.da_get(q);
    if (!m) return std::nullopt;
    if (m->check == p) return (int64_t)(-m->base);
    return std::nullopt;
}

// da.rs:155-182
std::optional<std::vector<std::pair<int64_t, size_t>>> da_common_prefix(const Dict& d, std::string_view keyword,
                                                                        Counters* ctr) {
    int32_t p = (int32_t)ROOT_ID;
    std::vector<std::pair<int64_t, size_t>> out;
    for (size_t i = 0; i < keyword.size(); i++) {
        unsigned char ch = (unsigned char)keyword[i];
        int32_t prev = p;
        p = d.da.at((size_t)prev).base + (int32_t)ch;
        if (ctr) ctr->probes++;
        const DaNode* n = d.da_get(p);
        if (!n || n->check != prev) break;
        if (ctr) ctr->probes_ok++;
        int32_t ahead = d.da[(size_t)p].base + TERMINATOR;
        const DaNode* a = d.da_get(ahead);
        if (a && a->check == p && a->base < 0) out.emplace_back((int64_t)(-a->base), i + 1);
    }
    if (out.empty()) return std::nullopt;
    return out;
}

// index.rs:40-53
std::optional<std::vector<std::pair<int64_t, size_t>>> index_common_prefix(const Dict& d, std::string_view input,
                                                                           Counters* ctr) {
    auto hits = da_common_prefix(d, input, ctr);
    if (!hits) return std::nullopt;
    std::vector<std::pair<int64_t, size_t>> results;
    for (auto& [id, len] : *hits) {
        auto it = d.dup.find(id);
        size_t dup = it == d.dup.end() ? 0 : it->second;
        for (size_t i = 0; i <= dup; i++) results.emplace_back(id + (int64_t)i, len);
    }
    return results;
}

// --- UTF-8 helpers standing in for Rust's str::chars()/len_utf8() on VALID UTF-8 ---------------
inline size_t utf8_len(unsigned char lead) { return lead < 0x80 ? 1 : lead < 0xE0 ? 2 : lead < 0xF0 ? 3 : 4; }
inline uint32_t utf8_decode(const char* s, size_t n) {
    const unsigned char* u = (const unsigned char*)s;
    switch (n) {
        case 1: return u[0];
        case 2: return ((u[0] & 0x1Fu) << 6) | (u[1] & 0x3Fu);
        case 3: return ((u[0] & 0x0Fu) << 12) | ((u[1] & 0x3Fu) << 6) | (u[2] & 0x3Fu);
        default: return ((u[0] & 0x07u) << 18) | ((u[1] & 0x3Fu) << 12) | ((u[2] & 0x3Fu) << 6) | (u[3] & 0x3Fu);
    }
}
inline size_t chars_count(std::string_view s) {
    size_t n = 0;
    for (unsigned char c : s) n += (c & 0xC0) != 0x80;
    return n;
}

enum class Kind : int32_t { Dummy = 0, Known = 1, Unknown = 2 };  // node.rs:16-24 / token.rs:4-8

struct Node {  // node.rs:7-24 flattened (Dummy has no id/surface)
    Kind kind;
    int64_t id;
    size_t byte_pos, char_pos;
    Morph morph;
    std::string surface;
};

struct Lattice {  // lattice.rs:6-10
    const Dict* dict;
    std::vector<Node> nodes;
    std::vector<std::vector<size_t>> edges;
    Counters* ctr;

    Lattice(const Dict* d, std::string_view input, Counters* c)  // lattice.rs:13-20
        : dict(d), edges(chars_count(input) + 2), ctr(c) {}

    void add_bos_node() {  // lattice.rs:156-164
        size_t idx = nodes.size();
        nodes.push_back(Node{Kind::Dummy, BOS_EOS_ID, 0, 0, Morph{0, 0, 0}, std::string()});
        edges.at(0).push_back(idx);
    }
    void add_eos_node(std::string_view input) {  // lattice.rs:165-175
        size_t idx = nodes.size();
        size_t byte_pos = input.size();
        size_t char_pos = chars_count(input);
        nodes.push_back(Node{Kind::Dummy, BOS_EOS_ID, byte_pos, char_pos, Morph{0, 0, 0}, std::string()});
        edges.at(char_pos + 1).push_back(idx);
    }
    void add_known_node(int64_t id, size_t byte_pos, size_t char_pos, std::string_view surface) {  // :177-188
        Node n{Kind::Known, id, byte_pos, char_pos, dict->morphs.at((size_t)(id - 1)), std::string(surface)};
        size_t idx = nodes.size();
        nodes.push_back(std::move(n));
        edges.at(char_pos + chars_count(surface)).push_back(idx);
    }
    void add_unknown_node(int64_t id, size_t byte_pos, size_t char_pos, std::string_view surface) {  // :190-201
        Node n{Kind::Unknown, id, byte_pos, char_pos, dict->unk_morphs.at((size_t)(id - 1)), std::string(surface)};
        size_t idx = nodes.size();
        nodes.push_back(std::move(n));
        edges.at(char_pos + chars_count(surface)).push_back(idx);
    }

    bool process_known_words(size_t byte_pos, size_t char_pos, std::string_view input) {  // lattice.rs:24-38
        std::string_view text = input.substr(byte_pos);
        auto hits = index_common_prefix(*dict, text, ctr);
        if (!hits) return false;
        for (auto& [id, byte_length] : *hits) {
            size_t end_byte_pos = byte_pos + byte_length;
            add_known_node(id, byte_pos, char_pos, input.substr(byte_pos, end_byte_pos - byte_pos));
        }
        return true;
    }

    void process_unknown_words(size_t byte_pos, size_t char_pos, uint32_t ch, size_t ch_len, std::string_view input,
                               bool matched_known) {  // lattice.rs:42-99
        uint8_t cat = dict->category(ch);
        if (!matched_known || dict->invoke_list.at(cat)) {
            constexpr size_t MAXIMUM_UNKNOWN_WORD_LENGTH = 1024;
            bool is_group = cat < dict->group_list.size() ? (bool)dict->group_list[cat] : false;
            size_t end_byte_pos = byte_pos + ch_len;
            size_t unknown_word_length = 1;
            if (is_group) {
                size_t q = end_byte_pos;
                while (q < input.size()) {
                    size_t l = utf8_len((unsigned char)input[q]);
                    uint32_t next_char = utf8_decode(input.data() + q, l);
                    if (dict->category(next_char) != cat) break;
                    end_byte_pos += l;
                    q += l;
                    unknown_word_length += 1;
                    if (unknown_word_length >= MAXIMUM_UNKNOWN_WORD_LENGTH) break;
                }
            }
            auto it = dict->unk_map.find(cat);
            if (it != dict->unk_map.end()) {
                int64_t morph_id = it->second.first;
                size_t count = it->second.second;
                std::string_view surface = input.substr(byte_pos, end_byte_pos - byte_pos);
                for (size_t i = 0; i < count; i++) add_unknown_node(morph_id + (int64_t)i, byte_pos, char_pos, surface);
            }
        }
    }

    static Lattice build(const Dict* d, std::string_view input, Counters* c) {  // lattice.rs:101-114
        size_t byte_pos = 0;
        Lattice la(d, input, c);
        la.add_bos_node();
        size_t char_pos = 0;
        while (byte_pos < input.size()) {
            size_t l = utf8_len((unsigned char)input[byte_pos]);
            uint32_t ch = utf8_decode(input.data() + byte_pos, l);
            bool matched_known = la.process_known_words(byte_pos, char_pos, input);
            la.process_unknown_words(byte_pos, char_pos, ch, l, input, matched_known);
            byte_pos += l;
            char_pos += 1;
        }
        la.add_eos_node(input);
        return la;
    }

    // lattice.rs:116-154.  Returns the path as node indices (the reference clones the nodes); dp/pre are
    // optionally exported for node-level parity.
    std::vector<size_t> viterbi(std::vector<std::optional<int32_t>>* dp_out = nullptr,
                                std::vector<std::optional<size_t>>* pre_out = nullptr) const {
        constexpr int32_t INF = 1 << 30;
        std::vector<std::optional<int32_t>> dp(nodes.size());
        std::vector<std::optional<size_t>> pre_nodes(nodes.size());
        size_t char_len = edges.size();
        for (size_t char_pos = 1; char_pos < char_len; char_pos++) {
            for (size_t i : edges[char_pos]) {
                const Node& target = nodes[i];
                dp[i] = INF;
                size_t tpos = target.char_pos;
                for (size_t j : edges[tpos]) {
                    const Node& previous = nodes[j];
                    int32_t prev_cost = dp[j].value_or(0);
                    int32_t cost = (int32_t)target.morph.cost;
                    int32_t matrix_cost =
                        (int32_t)dict->conn_get((size_t)previous.morph.right_id, (size_t)target.morph.left_id);
                    int32_t total_cost = std::min(prev_cost + cost + matrix_cost, INF);
                    if (ctr) ctr->pairs++;
                    if (!dp[i].has_value() || total_cost < *dp[i]) {
                        dp[i] = total_cost;
                        pre_nodes[i] = j;
                    }
                }
            }
        }
        size_t pos = nodes.size() - 1;
        std::vector<size_t> paths;
        while (pre_nodes[pos].has_value()) {
            paths.push_back(pos);
            pos = *pre_nodes[pos];
        }
        std::reverse(paths.begin(), paths.end());
        if (dp_out) *dp_out = std::move(dp);
        if (pre_out) *pre_out = std::move(pre_nodes);
        return paths;
    }
};

struct Token {  // token.rs:11-18 (surface is rebuilt by the caller from position + byte length)
    int64_t id;
    int64_t cls;
    int64_t position, start, end;
    int64_t byte_len;  // not in the reference struct: len(surface) in bytes, so callers can slice; 3 for "EOS"
};

// tokenizer.rs:16-45
std::vector<Token> tokenize(const Dict& d, std::string_view input, Counters* ctr, int32_t* eos_cost) {
    Lattice lattice = Lattice::build(&d, input, ctr);
    std::vector<std::optional<int32_t>> dp;
    std::vector<size_t> path = lattice.viterbi(&dp, nullptr);
    std::vector<Token> out;
    out.reserve(path.size());
    for (size_t idx : path) {
        const Node& node = lattice.nodes[idx];
        std::string surface = node.kind == Kind::Dummy ? std::string("EOS") : node.surface;  // clone, tokenizer.rs:28-31
        size_t char_pos = node.char_pos;
        size_t end_pos = char_pos + chars_count(surface);
        out.push_back(Token{node.id, (int64_t)node.kind, (int64_t)node.byte_pos, (int64_t)char_pos, (int64_t)end_pos,
                            (int64_t)surface.size()});
    }
    if (ctr) {
        ctr->bytes += input.size();
        ctr->chars += chars_count(input);
        ctr->nodes += lattice.nodes.size();
        ctr->tokens += out.size();
    }
    if (eos_cost) *eos_cost = dp.back().value_or(0);
    return out;
}

void add(Counters& a, const Counters& b) {
    a.bytes += b.bytes; a.chars += b.chars; a.probes += b.probes; a.probes_ok += b.probes_ok;
    a.nodes += b.nodes; a.pairs += b.pairs; a.tokens += b.tokens;
}

}  // namespace

extern "C" {

struct ko_dict;  // opaque = Dict

ko_dict* ko_dict_create(const int32_t* da, uint64_t da_len, const int64_t* dup_ids, const uint64_t* dup_counts,
                        uint64_t n_dup, const int16_t* morphs, uint64_t n_morphs, uint64_t conn_row, uint64_t conn_col,
                        const int16_t* conn, const uint8_t* char_category, uint64_t n_char_category,
                        const uint8_t* invoke, uint64_t n_invoke, const uint8_t* group, uint64_t n_group,
                        const uint8_t* unk_cat, const int64_t* unk_first, const uint64_t* unk_count, uint64_t n_unk_map,
                        const int16_t* unk_morphs, uint64_t n_unk_morphs) {
    Dict* d = new Dict();
    d->da.resize(da_len);
    for (uint64_t i = 0; i < da_len; i++) d->da[i] = DaNode{da[2 * i], da[2 * i + 1]};
    for (uint64_t i = 0; i < n_dup; i++) d->dup[dup_ids[i]] = (size_t)dup_counts[i];
    d->morphs.resize(n_morphs);
    for (uint64_t i = 0; i < n_morphs; i++) d->morphs[i] = Morph{morphs[3 * i], morphs[3 * i + 1], morphs[3 * i + 2]};
    d->conn_row = conn_row;
    d->conn_col = conn_col;
    d->conn.assign(conn, conn + conn_row * conn_col);
    d->char_category.assign(char_category, char_category + n_char_category);
    for (uint64_t i = 0; i < n_invoke; i++) d->invoke_list.push_back(invoke[i] != 0);
    for (uint64_t i = 0; i < n_group; i++) d->group_list.push_back(group[i] != 0);
    for (uint64_t i = 0; i < n_unk_map; i++) d->unk_map[unk_cat[i]] = {unk_first[i], (size_t)unk_count[i]};
    d->unk_morphs.resize(n_unk_morphs);
    for (uint64_t i = 0; i < n_unk_morphs; i++)
        d->unk_morphs[i] = Morph{unk_morphs[3 * i], unk_morphs[3 * i + 1], unk_morphs[3 * i + 2]};
    return (ko_dict*)d;
}

void ko_dict_destroy(ko_dict* d) { delete (Dict*)d; }

// da.rs:133-153; returns id or 0 for None.
int64_t ko_da_search(const ko_dict* d, const uint8_t* s, uint64_t n) {
    auto r = da_search(*(const Dict*)d, std::string_view((const char*)s, n));
    return r ? *r : 0;
}

// index.rs:40-53 (use_dup=1) or da.rs:155-182 (use_dup=0).  Writes up to cap (id,len) pairs; returns
// the number of results, or -1 for None.
int64_t ko_common_prefix(const ko_dict* d, const uint8_t* s, uint64_t n, int use_dup, int64_t* ids, uint64_t* lens,
                         uint64_t cap) {
    std::string_view sv((const char*)s, n);
    auto r = use_dup ? index_common_prefix(*(const Dict*)d, sv, nullptr) : da_common_prefix(*(const Dict*)d, sv, nullptr);
    if (!r) return -1;
    for (size_t i = 0; i < r->size() && i < cap; i++) {
        ids[i] = (*r)[i].first;
        lens[i] = (*r)[i].second;
    }
    return (int64_t)r->size();
}

int16_t ko_conn_get(const ko_dict* d, uint64_t row, uint64_t col) { return ((const Dict*)d)->conn_get(row, col); }

struct ko_result {
    uint64_t n_sent;
    uint64_t* tok_off;   // [n_sent+1]
    int64_t* tokens;     // [n_tok][6] = id, class, position, start, end, byte_len
    int32_t* eos_cost;   // [n_sent]
    uint64_t counters[7];  // B, C, P, P_ok, N, E, T
};

// Batched Tokenizer::tokenize over sentences [off[i], off[i+1]) of one blob.  n_threads>1 shards the
// sentences contiguously over std::threads sharing the read-only dict (legal: Dict is Send+Sync,
// SURVEY §8b).  collect=0 discards tokens (timing only: the per-call Vec<Token> is still built).
ko_result* ko_tokenize_batch(const ko_dict* dp, const uint8_t* utf8, const uint64_t* off, uint64_t n_sent, int n_threads,
                             int collect) {
    const Dict& d = *(const Dict*)dp;
    if (n_threads < 1) n_threads = 1;
    std::vector<std::vector<Token>> per(collect ? n_sent : 0);
    std::vector<int32_t> costs(n_sent);
    std::vector<Counters> ctrs((size_t)n_threads);
    std::vector<uint64_t> ntok(n_sent);
    auto work = [&](int t) {
        uint64_t lo = n_sent * (uint64_t)t / (uint64_t)n_threads, hi = n_sent * (uint64_t)(t + 1) / (uint64_t)n_threads;
        for (uint64_t s = lo; s < hi; s++) {
            std::string_view in((const char*)utf8 + off[s], off[s + 1] - off[s]);
            std::vector<Token> toks = tokenize(d, in, &ctrs[(size_t)t], &costs[s]);
            ntok[s] = toks.size();
            if (collect) per[s] = std::move(toks);
        }
    };
    if (n_threads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    ko_result* r = (ko_result*)calloc(1, sizeof(ko_result));
    r->n_sent = n_sent;
    r->tok_off = (uint64_t*)malloc((n_sent + 1) * sizeof(uint64_t));
    r->tok_off[0] = 0;
    for (uint64_t s = 0; s < n_sent; s++) r->tok_off[s + 1] = r->tok_off[s] + ntok[s];
    r->eos_cost = (int32_t*)malloc((n_sent ? n_sent : 1) * sizeof(int32_t));
    memcpy(r->eos_cost, costs.data(), n_sent * sizeof(int32_t));
    uint64_t total = r->tok_off[n_sent];
    r->tokens = (int64_t*)malloc((total ? total : 1) * 6 * sizeof(int64_t));
    if (collect) {
        for (uint64_t s = 0; s < n_sent; s++) {
            int64_t* o = r->tokens + r->tok_off[s] * 6;
            for (auto& t : per[s]) {
                o[0] = t.id; o[1] = t.cls; o[2] = t.position; o[3] = t.start; o[4] = t.end; o[5] = t.byte_len;
                o += 6;
            }
        }
    }
    Counters tot;
    for (auto& c : ctrs) add(tot, c);
    uint64_t cs[7] = {tot.bytes, tot.chars, tot.probes, tot.probes_ok, tot.nodes, tot.pairs, tot.tokens};
    memcpy(r->counters, cs, sizeof(cs));
    return r;
}

void ko_result_free(ko_result* r) {
    if (!r) return;
    free(r->tok_off);
    free(r->tokens);
    free(r->eos_cost);
    free(r);
}

struct ko_lattice {
    uint64_t n_nodes;
    int64_t* nodes;   // [n_nodes][9] = kind, id, byte_pos, char_pos, end_char_pos (edges bucket), left, right, cost, byte_len
    int64_t* dp;      // [n_nodes]  (None -> INT64_MIN)
    int64_t* pre;     // [n_nodes]  (None -> -1)
    uint64_t n_path;
    uint64_t* path;   // node indices of viterbi()
    uint64_t n_buckets;
    uint64_t* edge_off;  // [n_buckets+1]
    uint64_t* edge_idx;  // concatenated edges[]
};

// Lattice{nodes,edges} + viterbi() internals for node-level parity (lattice.rs:6-10,116-154).
ko_lattice* ko_lattice_dump(const ko_dict* dp_, const uint8_t* utf8, uint64_t n) {
    const Dict& d = *(const Dict*)dp_;
    std::string_view in((const char*)utf8, n);
    Lattice la = Lattice::build(&d, in, nullptr);
    std::vector<std::optional<int32_t>> dp;
    std::vector<std::optional<size_t>> pre;
    std::vector<size_t> path = la.viterbi(&dp, &pre);
    ko_lattice* r = (ko_lattice*)calloc(1, sizeof(ko_lattice));
    r->n_nodes = la.nodes.size();
    r->nodes = (int64_t*)malloc(r->n_nodes * 9 * sizeof(int64_t));
    r->dp = (int64_t*)malloc(r->n_nodes * sizeof(int64_t));
    r->pre = (int64_t*)malloc(r->n_nodes * sizeof(int64_t));
    std::vector<int64_t> endpos(la.nodes.size(), -1);
    for (size_t b = 0; b < la.edges.size(); b++)
        for (size_t i : la.edges[b]) endpos[i] = (int64_t)b;
    for (size_t i = 0; i < la.nodes.size(); i++) {
        const Node& nd = la.nodes[i];
        int64_t* o = r->nodes + i * 9;
        o[0] = (int64_t)nd.kind; o[1] = nd.id; o[2] = (int64_t)nd.byte_pos; o[3] = (int64_t)nd.char_pos; o[4] = endpos[i];
        o[5] = nd.morph.left_id; o[6] = nd.morph.right_id; o[7] = nd.morph.cost; o[8] = (int64_t)nd.surface.size();
        r->dp[i] = dp[i] ? (int64_t)*dp[i] : INT64_MIN;
        r->pre[i] = pre[i] ? (int64_t)*pre[i] : -1;
    }
    r->n_path = path.size();
    r->path = (uint64_t*)malloc((path.size() ? path.size() : 1) * sizeof(uint64_t));
    for (size_t i = 0; i < path.size(); i++) r->path[i] = path[i];
    r->n_buckets = la.edges.size();
    r->edge_off = (uint64_t*)malloc((r->n_buckets + 1) * sizeof(uint64_t));
    r->edge_off[0] = 0;
    for (size_t b = 0; b < la.edges.size(); b++) r->edge_off[b + 1] = r->edge_off[b] + la.edges[b].size();
    r->edge_idx = (uint64_t*)malloc((r->edge_off[r->n_buckets] ? r->edge_off[r->n_buckets] : 1) * sizeof(uint64_t));
    size_t k = 0;
    for (auto& e : la.edges)
        for (size_t i : e) r->edge_idx[k++] = i;
    return r;
}

void ko_lattice_free(ko_lattice* r) {
    if (!r) return;
    free(r->nodes); free(r->dp); free(r->pre); free(r->path); free(r->edge_off); free(r->edge_idx);
    free(r);
}

}  // extern "C"
