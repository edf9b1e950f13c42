// C++ rendition of the reference's integration tests (src/tests.rs:8-202) against the C++ host mirror
// (include/kanpyo_b200.hpp) of the Rust API.  Same fixture dictionary, same assertions, plus the exact
// tokens the oracle derives for the fixture.  Built and run by tests/test_gpu_cpp.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kanpyo_b200.hpp"

using kanpyo::Dict;
using kanpyo::Token;
using kanpyo::TokenClass;
using kanpyo::Tokenizer;

#define CHECK(cond, msg)                                                       \
    do {                                                                       \
        if (!(cond)) {                                                         \
            std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, msg); \
            std::exit(1);                                                      \
        }                                                                      \
    } while (0)

static size_t chars(const std::string& s) {
    size_t n = 0;
    for (unsigned char c : s) n += (c & 0xC0) != 0x80;
    return n;
}

// create_test_dict(), src/tests.rs:8-108
struct Fixture {
    std::vector<uint8_t> category;
    int32_t* da = nullptr;
    kp_dict_arrays arrays;
    ~Fixture() { kp_da_free(da); }
};

static void fixture_arrays(Fixture* fx) {
    static const std::vector<std::string> keywords = {"テスト", "辞書", "形態素"};   // already in byte order
    std::string blob;
    std::vector<uint64_t> off = {0};
    std::vector<int64_t> ids;
    for (size_t i = 0; i < keywords.size(); i++) {
        blob += keywords[i];
        off.push_back(blob.size());
        ids.push_back((int64_t)i + 1);
    }
    uint64_t da_len = 0;
    kanpyo::check(kp_da_build((const uint8_t*)blob.data(), off.data(), keywords.size(), ids.data(), &fx->da, &da_len));
    static const int16_t morphs[] = {0, 0, 1000, 1, 1, 1200, 2, 2, 1100};
    static const int16_t conn[] = {0, 100, 200, 100, 0, 100, 200, 100, 0};
    fx->category.assign(1 << 16, 0);
    for (uint32_t c = 0x3042; c <= 0x3093; c++) fx->category[c] = 2;   // 'あ'..='ん' HIRAGANA
    for (uint32_t c = 0x4E00; c <= 0x9FA5; c++) fx->category[c] = 1;   // '一'..='龥' KANJI
    static const uint8_t invoke[] = {0, 1, 1}, group[] = {0, 1, 1};
    static const uint8_t unk_cat[] = {1, 2};
    static const int64_t unk_first[] = {1, 2};
    static const uint64_t unk_count[] = {1, 1};
    static const int16_t unk_morphs[] = {0, 0, 5000, 1, 1, 5000};
    kp_dict_arrays& a = fx->arrays;
    std::memset(&a, 0, sizeof(a));
    a.da = fx->da;
    a.da_len = da_len;
    a.morphs = morphs;
    a.n_morphs = 3;
    a.conn_row = 3;
    a.conn_col = 3;
    a.conn = conn;
    a.char_category = fx->category.data();
    a.n_char_category = fx->category.size();
    a.invoke_list = invoke;
    a.n_invoke = 3;
    a.group_list = group;
    a.n_group = 3;
    a.unk_cat = unk_cat;
    a.unk_first_id = unk_first;
    a.unk_count = unk_count;
    a.n_unk_map = 2;
    a.unk_morphs = unk_morphs;
    a.n_unk_morphs = 2;
}

static void test_tokenizer_basic(Tokenizer& tokenizer) {           // src/tests.rs:111-129
    auto tokens = tokenizer.tokenize("テスト");
    CHECK(!tokens.empty(), "Should tokenize known word");
    size_t non_eos = 0;
    for (auto& t : tokens) non_eos += t.token_class != TokenClass::Dummy;
    CHECK(non_eos > 0, "Should have at least one non-EOS token");
    // exact tokens (oracle): Known id 1 (0,0,3) + EOS (9,3,6), path cost 1000
    auto [toks, cost] = tokenizer.tokenize_with_cost("テスト");
    CHECK(toks.size() == 2 && cost == 1000, "テスト -> 2 tokens, cost 1000");
    CHECK((toks[0] == Token{1, TokenClass::Known, 0, 0, 3, "テスト"}), "first token");
    CHECK((toks[1] == Token{0, TokenClass::Dummy, 9, 3, 6, "EOS"}), "EOS token");
}

static void test_tokenizer_empty_input(Tokenizer& tokenizer) {     // src/tests.rs:131-143
    auto tokens = tokenizer.tokenize("");
    CHECK(!tokens.empty(), "Should produce EOS token for empty input");
    CHECK((tokens[0] == Token{0, TokenClass::Dummy, 0, 0, 3, "EOS"}), "EOS of the empty input");
}

static void test_tokenizer_unknown_word(Tokenizer& tokenizer) {    // src/tests.rs:145-154
    auto [tokens, cost] = tokenizer.tokenize_with_cost("あいうえお");
    CHECK(!tokens.empty(), "Should tokenize unknown words");
    CHECK(tokens.size() == 2 && cost == 5200, "あいうえお -> Unknown + EOS, cost 5200");
    CHECK((tokens[0] == Token{2, TokenClass::Unknown, 0, 0, 5, "あいうえお"}), "unknown token");
    CHECK((tokens[1] == Token{0, TokenClass::Dummy, 15, 5, 8, "EOS"}), "EOS token");
}

static void test_token_positions(Tokenizer& tokenizer) {           // src/tests.rs:156-176
    const std::string input = "テスト";
    for (auto& token : tokenizer.tokenize(input)) {
        if (token.token_class != TokenClass::Dummy) {
            CHECK(token.start <= token.end, "Token start should be <= end");
            CHECK(token.end <= chars(input), "Token end should be within input");
            CHECK(token.length() == chars(token.surface), "length() == chars(surface)");
        }
    }
}

static void test_tokenizer_dict_roundtrip(Tokenizer& tokenizer1, int device) {   // src/tests.rs:178-202
    // the persistent form here is the packed blob (the reference's zip container is out of scope)
    const void* blob = nullptr;
    uint64_t size = 0;
    kanpyo::check(kp_dict_blob(tokenizer1.dict().handle(), &blob, &size));
    Tokenizer tokenizer2(Dict::from_blob(blob, size, device));
    auto tokens1 = tokenizer1.tokenize("テスト"), tokens2 = tokenizer2.tokenize("テスト");
    CHECK(tokens1.size() == tokens2.size(), "Should produce same number of tokens");
    for (size_t i = 0; i < tokens1.size(); i++) CHECK(tokens1[i] == tokens2[i], "Tokens should be equal");
}

static void test_batch_and_lattice(Tokenizer& tokenizer) {
    auto all = tokenizer.tokenize_batch({"テスト", "", "辞書テスト形態素", "あいうえお"});
    CHECK(all.size() == 4 && all[1].size() == 1 && all[0] == tokenizer.tokenize("テスト"), "batch == per sentence");
    CHECK(all[2].size() == 4 && all[2][1].surface == "テスト" && all[2][2].id == 3, "辞書/テスト/形態素/EOS");
    auto la = tokenizer.lattice("あいうえお");                     // BOS + 5 unknown nodes + EOS, four of them dead
    CHECK(la.size() == 7 && !la[0].has_dp && la[1].dp == 5100 && la[2].dp == (1 << 30) && la[2].pre == -1,
          "lattice dump");
    CHECK(la[6].dp == 5200 && la[6].pre == 1, "EOS dp / pre");
}

// ABI 2 from compiled host code: compact records + host expansion, both device paths, batches in flight, and the
// single-process shard group (one device here) must all give the tokens of the plain call -- including the empty
// paths of the fixture (no unknown entry for DEFAULT: 'x' has no node at all, EOS stays unreachable).
static void test_abi2_paths(Tokenizer& tokenizer, const kp_dict_arrays& arrays, int device) {
    const std::vector<std::string_view> inputs = {"テスト", "", "辞書テスト形態素", "あいうえお", "xテスト", "テストx", "x", "テxスト"};
    const auto want = tokenizer.tokenize_batch(inputs);
    CHECK(want[0].size() == 2 && want[2].size() == 4, "plain sentences");
    CHECK(want[4].empty() && want[5].empty() && want[6].empty() && want[7].empty(), "a char without any node: dp[EOS] = INF, no token at all");
    CHECK(tokenizer.tokenize_batch_compact(inputs) == want, "compact records == wide records");
    for (int path : {KP_PATH_PIPELINE, KP_PATH_FUSED, KP_PATH_AUTO}) {
        tokenizer.set_path(path);
        CHECK(tokenizer.tokenize_batch(inputs) == want, "device paths agree");
        CHECK(tokenizer.tokenize_batch_compact(inputs) == want, "device paths agree (compact)");
    }
    kanpyo::Queue queue(tokenizer.dict(), 2);
    std::vector<uint64_t> tickets;
    for (int k = 0; k < 2; k++) tickets.push_back(queue.submit(inputs));
    for (uint64_t t : tickets) CHECK(queue.wait(t) == want, "queued batch == plain call");
    kanpyo::Shards shards(arrays, {device});
    CHECK(shards.tokenize_batch(inputs) == want, "shard group == plain call");
}

static void test_errors(Tokenizer& tokenizer) {
    bool threw = false;
    try {
        tokenizer.tokenize(std::string_view("\xff\xfe", 2));
    } catch (const kanpyo::Error& e) {
        threw = e.status() == KP_ERR_UTF8;
    }
    CHECK(threw, "invalid UTF-8 must raise KP_ERR_UTF8");
}

int main(int argc, char** argv) {
    const int device = argc > 1 ? std::atoi(argv[1]) : 0;
    try {
        Fixture fx;
        fixture_arrays(&fx);
        Tokenizer tokenizer(Dict(fx.arrays, device));
        test_tokenizer_basic(tokenizer);
        test_tokenizer_empty_input(tokenizer);
        test_tokenizer_unknown_word(tokenizer);
        test_token_positions(tokenizer);
        test_tokenizer_dict_roundtrip(tokenizer, device);
        test_batch_and_lattice(tokenizer);
        test_abi2_paths(tokenizer, fx.arrays, device);
        test_errors(tokenizer);
    } catch (const kanpyo::Error& e) {
        std::fprintf(stderr, "kanpyo::Error %d: %s\n", e.status(), e.what());
        return e.status() == KP_ERR_CUDA ? 77 : 1;
    }
    std::puts("cpp tests ok");
    return 0;
}
