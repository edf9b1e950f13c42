// kanpyo_b200.hpp — C++17 host-side mirror of the reference's Rust API for the tokenizer path, over the
// C ABI of kanpyo_b200.h (header-only; link with -lkanpyo_b200).
//
//   reference (Rust)                                              here
//   kanpyo::token::TokenClass            src/token.rs:4-8          kanpyo::TokenClass
//   kanpyo::token::Token                 src/token.rs:11-18,39-53  kanpyo::Token  (length(), operator==)
//   kanpyo_dict::dict::Dict              kanpyo-dict/src/dict.rs:21-30   kanpyo::Dict  (device-resident, immutable)
//   kanpyo::tokenizer::Tokenizer::new    src/tokenizer.rs:12-14    kanpyo::Tokenizer(dict)
//   Tokenizer::tokenize(&self, &str)     src/tokenizer.rs:16-45    Tokenizer::tokenize(std::string_view)
//   kanpyo::lattice::node::Node          src/lattice/node.rs:16-52 kanpyo::LatticeNode
//   Lattice::build + Lattice::viterbi    src/lattice.rs:101,116    Tokenizer::lattice(std::string_view)
//   (new) batches in flight                                        kanpyo::Queue   (kp_queue_*: copies hidden behind kernels)
//   (new) every GPU of the box, one process                        kanpyo::Shards  (kp_shards_*: NCCL broadcast, sentence shards)
//
// Error behaviour: the reference's `tokenize` is infallible by signature and panics on an index out
// of range; here a non-zero C-ABI status (no GPU, invalid UTF-8, out of memory) throws kanpyo::Error,
// the C++ counterpart of that panic.  There is no CPU fallback.
//
// Threading: a Dict may be shared by any number of Tokenizers / threads (it is immutable, like the
// reference's `Dict`); one Tokenizer owns one CUDA stream and its scratch, so use one per thread
// (the reference's `tokenize(&self)` is re-entrant; clone the Tokenizer handle per thread here).
#ifndef KANPYO_B200_HPP
#define KANPYO_B200_HPP

#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

#include "kanpyo_b200.h"

namespace kanpyo {

class Error : public std::runtime_error {
  public:
    Error(int status, const std::string& what) : std::runtime_error(what), status_(status) {}
    int status() const { return status_; }

  private:
    int status_;
};

inline void check(int status) {
    if (status != KP_OK) {
        const char* detail = kp_last_error();
        throw Error(status, std::string(kp_strerror(status)) + (detail && *detail ? std::string(": ") + detail : ""));
    }
}

enum class TokenClass : uint8_t { Dummy = KP_CLASS_DUMMY, Known = KP_CLASS_KNOWN, Unknown = KP_CLASS_UNKNOWN };

struct Token {                      // src/token.rs:11-18
    int64_t id;                     // KeywordID (isize)
    TokenClass token_class;         // `class` is a keyword in C++
    size_t position;                // byte position
    size_t start;                   // char position
    size_t end;                     // char position
    std::string surface;
    size_t length() const { return end - start; }                       // src/token.rs:39-41
    bool operator==(const Token& o) const {                              // src/token.rs:44-53
        return id == o.id && token_class == o.token_class && position == o.position && start == o.start &&
               end == o.end && surface == o.surface;
    }
    bool operator!=(const Token& o) const { return !(*this == o); }
};

struct LatticeNode {                // Node + viterbi() internals, src/lattice/node.rs:16-52, lattice.rs:116-143
    int64_t id;
    TokenClass node_class;
    size_t byte_pos, char_pos, end_char_pos;
    int16_t left_id, right_id, cost;
    bool has_dp;                    // false where the reference holds None (BOS)
    int32_t dp;
    int64_t pre;                    // -1 = None
};

// Read-only dictionary staged in HBM once.  Build it from the flat arrays the reference's `Dict`
// holds, or from a packed blob (kp_dict_pack / Dict.pack() in the Python host API).
class Dict {
  public:
    Dict(const kp_dict_arrays& arrays, int device = 0) {
        kp_dict* d = nullptr;
        check(kp_dict_create(&arrays, device, &d));
        h_.reset(d, kp_dict_destroy);
    }
    static Dict from_blob(const void* blob, uint64_t size, int device = 0) {
        kp_dict* d = nullptr;
        check(kp_dict_create_from_blob(blob, size, device, &d));
        return Dict(d);
    }
    static Dict load_blob(const std::string& path, int device = 0) {
        std::FILE* f = std::fopen(path.c_str(), "rb");
        if (!f) throw Error(KP_ERR_ARG, "cannot open " + path);
        std::vector<char> buf;
        char tmp[1 << 16];
        size_t n;
        while ((n = std::fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
        std::fclose(f);
        return from_blob(buf.data(), buf.size(), device);
    }
    const kp_dict* handle() const { return h_.get(); }

  private:
    explicit Dict(kp_dict* d) : h_(d, kp_dict_destroy) {}
    std::shared_ptr<kp_dict> h_;
};

class Tokenizer {
  public:
    explicit Tokenizer(Dict dict) : dict_(std::move(dict)) {            // Tokenizer::new, src/tokenizer.rs:12-14
        kp_tokenizer* t = nullptr;
        check(kp_tokenizer_create(dict_.handle(), &t));
        h_.reset(t, kp_tokenizer_destroy);
    }

    // Tokenizer::tokenize (src/tokenizer.rs:16-45): best path without BOS, with the EOS token
    // {id 0, Dummy, position = len bytes, start = n chars, end = n + 3, surface "EOS"}.
    std::vector<Token> tokenize(std::string_view input) const {
        kp_result r;
        check(kp_tokenize(h_.get(), reinterpret_cast<const uint8_t*>(input.data()), input.size(), &r));
        return materialize(input, r, 0);
    }

    // The same call over many independent sentences in one device pass.
    std::vector<std::vector<Token>> tokenize_batch(const std::vector<std::string_view>& inputs) const {
        std::string text;
        std::vector<uint64_t> off(inputs.size() + 1, 0);
        for (size_t i = 0; i < inputs.size(); i++) {
            text.append(inputs[i]);
            off[i + 1] = text.size();
        }
        kp_result r;
        check(kp_tokenize_batch(h_.get(), reinterpret_cast<const uint8_t*>(text.data()), off.data(), inputs.size(), &r));
        std::vector<std::vector<Token>> out(inputs.size());
        for (size_t i = 0; i < inputs.size(); i++) out[i] = materialize(inputs[i], r, i);
        return out;
    }

    // The same through the compact transfer records (kp_token8: half the device-to-host bytes); position / start
    // are rebuilt on the host by kp_expand_tokens8.
    std::vector<std::vector<Token>> tokenize_batch_compact(const std::vector<std::string_view>& inputs) const {
        std::string text;
        std::vector<uint64_t> off;
        pack(inputs, &text, &off);
        kp_result8 r;
        check(kp_tokenize_batch8(h_.get(), reinterpret_cast<const uint8_t*>(text.data()), off.data(), inputs.size(), &r));
        return materialize8(inputs, r, off);
    }

    // KP_PATH_AUTO / KP_PATH_PIPELINE / KP_PATH_FUSED: which device path serves the calls (identical results).
    void set_path(int path) { check(kp_tokenizer_set_path(h_.get(), path)); }

    static void pack(const std::vector<std::string_view>& inputs, std::string* text, std::vector<uint64_t>* off) {
        off->assign(inputs.size() + 1, 0);
        for (size_t i = 0; i < inputs.size(); i++) {
            text->append(inputs[i]);
            (*off)[i + 1] = text->size();
        }
    }

    static std::vector<std::vector<Token>> materialize8(const std::vector<std::string_view>& inputs, const kp_result8& r,
                                                        const std::vector<uint64_t>& off) {
        std::vector<kp_token> full(r.n_tokens);
        check(kp_expand_tokens8(&r, off.data(), full.data()));
        std::vector<uint64_t> tok_off(r.tok_off, r.tok_off + r.n_sent + 1);
        kp_result wide;
        wide.n_sent = r.n_sent;
        wide.n_tokens = r.n_tokens;
        wide.tok_off = tok_off.data();
        wide.tokens = full.data();
        wide.eos_cost = r.eos_cost;
        std::vector<std::vector<Token>> out(inputs.size());
        for (size_t i = 0; i < inputs.size(); i++) out[i] = materialize(inputs[i], wide, i);
        return out;
    }

    // dp[EOS] of the last tokenize() call's sentence `s` is not kept here; use tokenize_with_cost.
    std::pair<std::vector<Token>, int32_t> tokenize_with_cost(std::string_view input) const {
        kp_result r;
        check(kp_tokenize(h_.get(), reinterpret_cast<const uint8_t*>(input.data()), input.size(), &r));
        return {materialize(input, r, 0), r.eos_cost[0]};
    }

    // Lattice::build(&dict, input).nodes in insertion order (BOS first, EOS last) with the dp / pre_nodes
    // tables of Lattice::viterbi (src/lattice.rs:6-10,101-143).
    std::vector<LatticeNode> lattice(std::string_view input) const {
        kp_lattice la;
        check(kp_lattice_dump(h_.get(), reinterpret_cast<const uint8_t*>(input.data()), input.size(), &la));
        std::vector<LatticeNode> out(la.n_nodes);
        for (uint64_t i = 0; i < la.n_nodes; i++) {
            const kp_lattice_node& n = la.nodes[i];
            out[i] = LatticeNode{n.id,       static_cast<TokenClass>(n.cls), n.byte_pos, n.char_pos, n.end_char,
                                 n.left_id, n.right_id,                     n.cost,     n.dp != INT32_MIN,
                                 n.dp,      n.pre};
        }
        return out;
    }

    kp_tokenizer* handle() const { return h_.get(); }
    const Dict& dict() const { return dict_; }                          // `pub dict: Dict`, src/tokenizer.rs:8

    static std::vector<Token> materialize(std::string_view input, const kp_result& r, uint64_t s) {
        std::vector<Token> out;
        const uint64_t a = r.tok_off[s], b = r.tok_off[s + 1];
        out.reserve(b - a);
        for (uint64_t k = a; k < b; k++) {
            const kp_token& t = r.tokens[k];
            Token o;
            o.id = t.id;
            o.token_class = static_cast<TokenClass>(t.cls);
            o.position = t.position;
            o.start = t.start;
            o.end = static_cast<size_t>(t.start) + t.char_len;
            if (o.token_class == TokenClass::Dummy) o.surface = "EOS";                   // src/tokenizer.rs:28-29
            else o.surface = std::string(input.substr(t.position, r.tokens[k + 1].position - t.position));
            out.push_back(std::move(o));
        }
        return out;
    }

  private:
    Dict dict_;
    std::shared_ptr<kp_tokenizer> h_;
};

// Successive batches in flight on `depth` tokenizer contexts (kp_queue_*): submit() returns at once, wait() blocks
// for that batch's tokens.  The inputs are copied into the ticket, so the caller's strings may go away.
class Queue {
  public:
    explicit Queue(const Dict& dict, uint32_t depth = 3) : dict_(dict), depth_(depth), slots_(depth) {
        kp_queue* q = nullptr;
        check(kp_queue_create(dict_.handle(), depth, &q));
        h_.reset(q, kp_queue_destroy);
    }
    uint64_t submit(const std::vector<std::string_view>& inputs) {
        Slot& s = slots_[next_ % depth_];
        s.inputs.assign(inputs.begin(), inputs.end());
        s.text.clear();
        std::vector<std::string_view> views(s.inputs.begin(), s.inputs.end());
        Tokenizer::pack(views, &s.text, &s.off);
        uint64_t ticket = 0;
        check(kp_queue_submit(h_.get(), reinterpret_cast<const uint8_t*>(s.text.data()), s.off.data(), inputs.size(), &ticket));
        next_ = ticket + 1;
        return ticket;
    }
    std::vector<std::vector<Token>> wait(uint64_t ticket) {
        kp_result8 r;
        check(kp_queue_wait(h_.get(), ticket, &r));
        const Slot& s = slots_[ticket % depth_];
        std::vector<std::string_view> views(s.inputs.begin(), s.inputs.end());
        return Tokenizer::materialize8(views, r, s.off);
    }

  private:
    struct Slot {
        std::vector<std::string> inputs;
        std::string text;
        std::vector<uint64_t> off;
    };
    Dict dict_;
    uint32_t depth_;
    uint64_t next_ = 0;
    std::vector<Slot> slots_;
    std::shared_ptr<kp_queue> h_;
};

// Every GPU of the box from one process (kp_shards_*): the dictionary is packed once and broadcast with NCCL, a
// batch is split into byte-balanced contiguous sentence ranges, one per GPU.
class Shards {
  public:
    explicit Shards(const kp_dict_arrays& arrays, const std::vector<int>& devices = {}) {
        int n = (int)devices.size();
        if (n == 0) check(kp_device_count(&n));
        kp_shards* g = nullptr;
        check(kp_shards_create(&arrays, devices.empty() ? nullptr : devices.data(), n, &g));
        h_.reset(g, kp_shards_destroy);
    }
    std::vector<std::vector<Token>> tokenize_batch(const std::vector<std::string_view>& inputs) const {
        std::string text;
        std::vector<uint64_t> off;
        Tokenizer::pack(inputs, &text, &off);
        kp_result8 r;
        check(kp_shards_tokenize(h_.get(), reinterpret_cast<const uint8_t*>(text.data()), off.data(), inputs.size(), &r));
        return Tokenizer::materialize8(inputs, r, off);
    }

  private:
    std::shared_ptr<kp_shards> h_;
};

}  // namespace kanpyo
#endif  // KANPYO_B200_HPP
