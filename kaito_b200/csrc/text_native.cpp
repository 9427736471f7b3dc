// text_native.cpp -- host-side text analysis of the /index and /retrieve paths in native code:
//   * the bm25s analysis chain the reference reaches through BM25Retriever.from_defaults (hybrid_retriever.py:122-125):
//     lower-case, r"(?u)\b\w\w+\b" tokens, bm25s's 33 English stop words, Snowball English ("Porter2") stems -- PyStemmer
//     is C in the reference's stack too;
//   * BertTokenizer (uncased) as sentence-transformers runs it for bge (huggingface_local_embedding.py:34-53): BasicTokenizer
//     clean-up / lower-casing / punctuation split + greedy longest-match WordPiece -- HF tokenizers is native (Rust) there.
// Contract: the input is ASCII.  kaito_b200/text.py holds the Unicode-complete Python restatement, which is also the spec
// these functions are tested against (tests/test_text_native.py); the host calls the native path for ASCII text only.
// No CUDA in this file; it is part of libkaito_rag.so so that a cgo host gets the same analysis as the Python host.
#include <stdint.h>
#include <string.h>

#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/kaito_rag.h"

namespace {

// ------------------------------------------------------------------------------------------ Snowball English
inline bool is_vowel(char c) { return c == 'a' || c == 'e' || c == 'i' || c == 'o' || c == 'u' || c == 'y'; }
inline bool ends_with(const std::string& w, const char* suf)
{
    const size_t n = strlen(suf);
    return w.size() >= n && memcmp(w.data() + w.size() - n, suf, n) == 0;
}
inline bool starts_with(const std::string& w, const char* pre)
{
    const size_t n = strlen(pre);
    return w.size() >= n && memcmp(w.data(), pre, n) == 0;
}
size_t region_after_vc(const std::string& w, size_t start)
{
    for (size_t i = start + 1; i < w.size(); ++i)
        if (!is_vowel(w[i]) && is_vowel(w[i - 1])) return i + 1;
    return w.size();
}
bool contains_vowel(const std::string& w, size_t n)   // over w[0, n)
{
    for (size_t i = 0; i < n && i < w.size(); ++i) if (is_vowel(w[i])) return true;
    return false;
}
bool ends_short_syllable(const std::string& w)
{
    const size_t n = w.size();
    if (n == 2) return is_vowel(w[0]) && !is_vowel(w[1]);
    if (n >= 3) {
        const char l = w[n - 1];
        return !is_vowel(w[n - 3]) && is_vowel(w[n - 2]) && !is_vowel(l) && l != 'w' && l != 'x' && l != 'Y';
    }
    return false;
}
bool ends_with_double(const std::string& w)
{
    static const char* D[] = {"bb", "dd", "ff", "gg", "mm", "nn", "pp", "rr", "tt"};
    for (const char* d : D) if (ends_with(w, d)) return true;
    return false;
}
std::string unmark(std::string w)
{
    for (char& c : w) if (c == 'Y') c = 'y';
    return w;
}

const std::unordered_map<std::string, std::string>& exceptions()
{
    static const std::unordered_map<std::string, std::string> m = {
        {"skis", "ski"}, {"skies", "sky"}, {"dying", "die"}, {"lying", "lie"}, {"tying", "tie"}, {"idly", "idl"},
        {"gently", "gentl"}, {"ugly", "ugli"}, {"early", "earli"}, {"only", "onli"}, {"singly", "singl"}, {"sky", "sky"},
        {"news", "news"}, {"howe", "howe"}, {"atlas", "atlas"}, {"cosmos", "cosmos"}, {"bias", "bias"}, {"andes", "andes"}};
    return m;
}
const std::unordered_set<std::string>& exceptions_1a()
{
    static const std::unordered_set<std::string> s = {"inning", "outing", "canning", "herring", "earring", "proceed", "exceed", "succeed"};
    return s;
}

std::string stem(const std::string& word)
{
    if (word.size() <= 2) return word;
    auto ex = exceptions().find(word);
    if (ex != exceptions().end()) return ex->second;
    std::string w = word[0] == '\'' ? word.substr(1) : word;
    for (size_t i = 0; i < w.size(); ++i)
        if (w[i] == 'y' && (i == 0 || is_vowel(w[i - 1]))) w[i] = 'Y';
    size_t r1;
    if (starts_with(w, "gener") || starts_with(w, "arsen")) r1 = 5;
    else if (starts_with(w, "commun")) r1 = 6;
    else r1 = region_after_vc(w, 0);
    const size_t r2 = region_after_vc(w, r1);

    // step 0
    if (ends_with(w, "'s'")) w.resize(w.size() - 3);
    else if (ends_with(w, "'s")) w.resize(w.size() - 2);
    else if (ends_with(w, "'")) w.resize(w.size() - 1);
    // step 1a
    if (ends_with(w, "sses")) w.resize(w.size() - 2);
    else if (ends_with(w, "ied") || ends_with(w, "ies")) w.resize(w.size() > 4 ? w.size() - 2 : w.size() - 1);
    else if (ends_with(w, "us") || ends_with(w, "ss")) { /* unchanged */ }
    else if (ends_with(w, "s")) { if (w.size() >= 2 && contains_vowel(w, w.size() - 2)) w.resize(w.size() - 1); }
    if (exceptions_1a().count(w)) return unmark(w);
    // step 1b
    if (ends_with(w, "eedly")) { if (w.size() - 5 >= r1) w.resize(w.size() - 3); }
    else if (ends_with(w, "eed")) { if (w.size() - 3 >= r1) w.resize(w.size() - 1); }
    else {
        static const char* S[] = {"ingly", "edly", "ing", "ed"};
        for (const char* suf : S) {
            if (!ends_with(w, suf)) continue;
            const size_t n = w.size() - strlen(suf);
            if (contains_vowel(w, n)) {
                w.resize(n);
                if (ends_with(w, "at") || ends_with(w, "bl") || ends_with(w, "iz")) w += 'e';
                else if (ends_with_double(w)) w.resize(w.size() - 1);
                else if (ends_short_syllable(w) && r1 >= w.size()) w += 'e';
            }
            break;
        }
    }
    // step 1c
    if (w.size() > 2 && (w.back() == 'y' || w.back() == 'Y') && !is_vowel(w[w.size() - 2])) w.back() = 'i';
    // step 2
    {
        static const char* S[][2] = {{"ization", "ize"}, {"ational", "ate"}, {"fulness", "ful"}, {"ousness", "ous"}, {"iveness", "ive"},
                                     {"tional", "tion"}, {"biliti", "ble"}, {"lessli", "less"}, {"entli", "ent"}, {"ation", "ate"},
                                     {"alism", "al"}, {"aliti", "al"}, {"ousli", "ous"}, {"iviti", "ive"}, {"fulli", "ful"}, {"enci", "ence"},
                                     {"anci", "ance"}, {"abli", "able"}, {"izer", "ize"}, {"ator", "ate"}, {"alli", "al"}, {"bli", "ble"},
                                     {"ogi", "og"}, {"li", ""}};
        for (auto& p : S) {
            if (!ends_with(w, p[0])) continue;
            const size_t ls = strlen(p[0]);
            if (w.size() - ls >= r1) {
                if (strcmp(p[0], "ogi") == 0) {
                    if (w.size() > ls && w[w.size() - ls - 1] == 'l') { w.resize(w.size() - ls); w += p[1]; }
                } else if (strcmp(p[0], "li") == 0) {
                    if (w.size() > 2 && strchr("cdeghkmnrt", w[w.size() - 3]) != nullptr) w.resize(w.size() - 2);
                } else {
                    w.resize(w.size() - ls); w += p[1];
                }
            }
            break;
        }
    }
    // step 3
    {
        static const char* S[][2] = {{"ational", "ate"}, {"tional", "tion"}, {"alize", "al"}, {"icate", "ic"}, {"iciti", "ic"}, {"ative", ""},
                                     {"ical", "ic"}, {"ness", ""}, {"ful", ""}};
        for (auto& p : S) {
            if (!ends_with(w, p[0])) continue;
            const size_t ls = strlen(p[0]);
            if (w.size() - ls >= r1) {
                if (strcmp(p[0], "ative") == 0) { if (w.size() - ls >= r2) w.resize(w.size() - ls); }
                else { w.resize(w.size() - ls); w += p[1]; }
            }
            break;
        }
    }
    // step 4
    {
        static const char* S[] = {"ement", "ance", "ence", "able", "ible", "ment", "ant", "ent", "ism", "ate", "iti", "ous", "ive", "ize",
                                  "ion", "al", "er", "ic"};
        for (const char* suf : S) {
            if (!ends_with(w, suf)) continue;
            const size_t ls = strlen(suf);
            if (w.size() - ls >= r2) {
                if (strcmp(suf, "ion") == 0) { if (w.size() > 3 && (w[w.size() - 4] == 's' || w[w.size() - 4] == 't')) w.resize(w.size() - 3); }
                else w.resize(w.size() - ls);
            }
            break;
        }
    }
    // step 5
    if (ends_with(w, "e")) {
        const std::string head = w.substr(0, w.size() - 1);
        if (w.size() - 1 >= r2 || (w.size() - 1 >= r1 && !ends_short_syllable(head))) w.resize(w.size() - 1);
    } else if (ends_with(w, "l")) {
        if (w.size() - 1 >= r2 && w.size() >= 2 && w[w.size() - 2] == 'l') w.resize(w.size() - 1);
    }
    return unmark(w);
}

const std::unordered_set<std::string>& stopwords()
{
    static const std::unordered_set<std::string> s = {
        "a", "an", "and", "are", "as", "at", "be", "but", "by", "for", "if", "in", "into", "is", "it", "no", "not", "of",
        "on", "or", "such", "that", "the", "their", "then", "there", "these", "they", "this", "to", "was", "will", "with"};
    return s;
}
inline bool is_word_char(unsigned char c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_'; }
inline char lower(char c) { return (c >= 'A' && c <= 'Z') ? (char)(c + 32) : c; }

// stems of one ASCII text, '\n'-separated
void analyze(const char* text, int64_t len, std::string& out, int32_t& n)
{
    n = 0;
    std::string tok;
    for (int64_t i = 0; i <= len; ++i) {
        const bool wc = i < len && is_word_char((unsigned char)text[i]);
        if (wc) { tok += lower(text[i]); continue; }
        if (tok.size() >= 2 && !stopwords().count(tok)) {
            if (n) out += '\n';
            out += stem(tok);
            ++n;
        }
        tok.clear();
    }
}

// ------------------------------------------------------------------------------------------ WordPiece
inline bool is_ascii_punct(unsigned char c) { return (c >= 33 && c <= 47) || (c >= 58 && c <= 64) || (c >= 91 && c <= 96) || (c >= 123 && c <= 126); }

}  // namespace

struct krag_wordpiece {
    std::unordered_map<std::string, int32_t> vocab;
    int32_t unk = -1, cls = -1, sep = -1;
    bool lower_case = true;

    void piece(const std::string& word, std::vector<int32_t>& ids) const
    {
        if (word.size() > 100) { ids.push_back(unk); return; }
        const size_t mark = ids.size();
        size_t start = 0;
        std::string sub;
        while (start < word.size()) {
            size_t end = word.size();
            int32_t cur = -1;
            while (start < end) {
                sub.assign(start == 0 ? "" : "##");
                sub.append(word, start, end - start);
                auto it = vocab.find(sub);
                if (it != vocab.end()) { cur = it->second; break; }
                --end;
            }
            if (cur < 0) { ids.resize(mark); ids.push_back(unk); return; }
            ids.push_back(cur);
            start = end;
        }
    }
    // BasicTokenizer on ASCII + WordPiece; [CLS] ids[:max_len-2] [SEP]
    void encode(const char* text, int64_t len, int32_t max_len, std::vector<int32_t>& out) const
    {
        std::vector<int32_t> ids;
        std::string word;
        auto flush = [&] { if (!word.empty()) { piece(word, ids); word.clear(); } };
        for (int64_t i = 0; i < len; ++i) {
            const unsigned char c = (unsigned char)text[i];
            if (c == ' ' || c == '\t' || c == '\n' || c == '\r') { flush(); continue; }
            if (c < 0x20 || c == 0x7F) continue;                       // other control characters are dropped
            if (is_ascii_punct(c)) { flush(); word.assign(1, (char)c); flush(); continue; }
            word += lower_case ? lower((char)c) : (char)c;
        }
        flush();
        out.clear();
        out.push_back(cls);
        const size_t keep = max_len > 2 ? (size_t)(max_len - 2) : 0;
        out.insert(out.end(), ids.begin(), ids.begin() + (ids.size() < keep ? ids.size() : keep));
        out.push_back(sep);
    }
};

namespace {
thread_local std::string g_text_err;
template <class F>
int32_t text_guard(F&& f)
{
    try { f(); return KRAG_OK; }
    catch (const std::bad_alloc&) { return KRAG_E_OOM; }
    catch (const std::exception&) { return KRAG_E_INVALID; }
}
template <class F>
void parallel_for(int64_t n, int64_t total_bytes, F&& body)
{
    unsigned hw = std::thread::hardware_concurrency();
    int64_t nt = hw ? hw : 4;
    if (nt > n / 8) nt = n / 8;          // a thread per >= 8 texts ...
    if (nt > total_bytes / 32768) nt = total_bytes / 32768;   // ... and per >= 32 KB of text: a batch of queries is not worth a thread
    if (nt <= 1) { for (int64_t i = 0; i < n; ++i) body(i); return; }
    std::vector<std::thread> th;
    for (int64_t t = 0; t < nt; ++t)
        th.emplace_back([&, t] { for (int64_t i = t; i < n; i += nt) body(i); });
    for (auto& x : th) x.join();
}
}  // namespace

extern "C" {

int32_t krag_text_analyze(const char* text, int64_t len, char* out, int64_t cap, int64_t* out_len, int32_t* n_terms)
{
    return text_guard([&] {
        if (!text || !out_len || !n_terms || len < 0) throw std::invalid_argument("null argument");
        std::string s;
        int32_t n = 0;
        analyze(text, len, s, n);
        *out_len = (int64_t)s.size();
        *n_terms = n;
        if (out && (int64_t)s.size() <= cap) memcpy(out, s.data(), s.size());
    });
}

int32_t krag_wordpiece_create(const char* vocab, int64_t len, int32_t lower_case, krag_wordpiece** out)
{
    return text_guard([&] {
        if (!vocab || !out || len < 0) throw std::invalid_argument("null argument");
        krag_wordpiece* w = new krag_wordpiece();
        w->lower_case = lower_case != 0;
        // one token per line, exactly like "\n".join(tokens).split("\n"); a repeated token keeps its LAST id (dict semantics)
        int32_t id = 0;
        const char* p = vocab; const char* e = vocab + len;
        for (;;) {
            const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
            const char* q = nl ? nl : e;
            w->vocab[std::string(p, q)] = id++;
            if (!nl) break;
            p = nl + 1;
        }
        auto need = [&](const char* t) { auto it = w->vocab.find(t); if (it == w->vocab.end()) { delete w; throw std::invalid_argument(t); } return it->second; };
        w->unk = need("[UNK]"); w->cls = need("[CLS]"); w->sep = need("[SEP]");
        *out = w;
    });
}

int32_t krag_wordpiece_encode_batch(const krag_wordpiece* w, int64_t n, const char* texts, const int64_t* offsets /*[n+1]*/,
                                    int32_t max_len, int32_t* out_ids /*[n * max_len]*/, int32_t* out_n /*[n]*/)
{
    return text_guard([&] {
        if (!w || !texts || !offsets || !out_ids || !out_n || n < 0 || max_len < 2) throw std::invalid_argument("bad argument");
        parallel_for(n, offsets[n] - offsets[0], [&](int64_t i) {
            std::vector<int32_t> ids;
            w->encode(texts + offsets[i], offsets[i + 1] - offsets[i], max_len, ids);
            out_n[i] = (int32_t)ids.size();
            memcpy(out_ids + i * (int64_t)max_len, ids.data(), ids.size() * sizeof(int32_t));
        });
    });
}

int32_t krag_wordpiece_destroy(krag_wordpiece* w)
{
    delete w;
    return KRAG_OK;
}

}  // extern "C"
