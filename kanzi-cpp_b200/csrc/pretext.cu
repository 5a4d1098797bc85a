// pretext.cu -- HOST stage TEXT of the block pipeline: kanzi's dictionary text codec, both wire
// variants.  The variant follows the entropy coder of the stream (TransformFactory.hpp:227-242: variant 2 under
// NONE / HUFFMAN / ANS0 / RANGE, variant 1 otherwise), so `-l 3` and `-l 5` use variant 2 and `-l 6` variant 1.
//
//   statistics / text detection   TextCodec::computeStats, detectType   transform/TextCodec.cpp:217-428
//   static dictionary             TextCodec::createDictionary           :182-214   (words: text_dict.inc)
//   variant 1 (escape tokens)     TextCodec1::forward / inverse         :588-754, :856-1015
//   variant 2 (high-bit indexes)  TextCodec2::forward / inverse         :1076-1244, :1374-1581
//
// A block is scanned once: letters extend the current word (two rolling hashes, as written and with the case of
// the first letter flipped), a delimiter ends it; a word found in the dictionary (1024 static English words +
// up to 2^19 words of the block itself, learnt in order of appearance, oldest replaced first) is replaced by
// its index.  The decoder learns the same words from the literal text in the same order.  Every table decision
// (what is inserted, evicted, re-mapped when the word list grows) is part of the format and follows the
// reference operation by operation; the code is organised around one dictionary class used by both variants and
// both directions.  Host code only.
#include <string.h>
#include <vector>

#include "pre.h"

namespace {

const u32 MULT1 = 0x7FEB352Du, MULT2 = 0x846CA68Bu;
const int MAX_WORD = 31, STATIC_WORDS = 1024, MAX_WORDS = 1 << 19;
const u8 TOKEN_WORD = 0x0F, TOKEN_FLIP = 0x0E; // variant 1: index follows / index follows, first letter's case flipped
const u8 MODE_NOT_TEXT = 0x80, MODE_CRLF = 0x40, MODE_XML = 0x20;

const char STATIC_TEXT[] =
#include "text_dict.inc"
    ;

inline bool is_letter(u8 c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }

// 0: letter, 1: delimiter, -1: neither (TextCodec::init, :152-179)
inline int char_class(u8 c)
{
    if (is_letter(c))
        return 0;
    if ((c >= ' ' && c <= '/') || (c >= ':' && c <= '?'))
        return 1;
    switch (c) {
    case '\n': case '\r': case '\t': case '_': case '|': case '{': case '}': case '[': case ']':
        return 1;
    default:
        return -1;
    }
}

inline u32 hash_step(u32 h, u8 c) { return h * MULT1 ^ (u32)c * MULT2; }

struct Word {
    const u8* text; // where the word's letters are (static list, input block, or coded block on the decode side)
    u32 hash;
    int lenIdx;     // length << 24 | index
};

// The word list and its hash map, with the reference's exact update rules.
struct Dictionary {
    std::vector<Word> list;
    std::vector<int> slot; // hash & mask -> index into list, -1 = empty
    u32 mask;
    int fixedWords; // static words (+ the two escape entries of variant 1): never replaced
    int next;       // list position the next learnt word takes
    u8 escapes[2];

    void init(int count, int logSlots, bool withEscapes)
    {
        int lg = 13;
        if (count >= 1024) {
            lg = 31 - __builtin_clz((u32)(count / 128));
            lg = (lg > 18) ? 18 : (lg < 13 ? 13 : lg);
        }
        fixedWords = STATIC_WORDS + (withEscapes ? 2 : 0);
        const int size = (fixedWords > (1 << lg)) ? fixedWords : (1 << lg);
        list.resize((size_t)size);
        mask = (1u << logSlots) - 1u;
        slot.assign((size_t)1 << logSlots, -1);
        // static words: lower case, hashed as such
        const u8* p = reinterpret_cast<const u8*>(STATIC_TEXT);
        int w = 0;
        while (*p && w < STATIC_WORDS) {
            int len = 0;
            u32 h = MULT1;
            while (p[len] && p[len] != ' ')
                h = hash_step(h, p[len++]);
            list[w] = Word{ p, h, (len << 24) | w };
            w++;
            p += len;
            while (*p == ' ')
                p++;
        }
        if (withEscapes) { // variant 1 reaches its two tokens through dictionary entries of length 1
            escapes[0] = TOKEN_FLIP;
            escapes[1] = TOKEN_WORD;
            list[STATIC_WORDS] = Word{ &escapes[0], 0, (1 << 24) | STATIC_WORDS };
            list[STATIC_WORDS + 1] = Word{ &escapes[1], 0, (1 << 24) | (STATIC_WORDS + 1) };
        }
        for (int i = 0; i < fixedWords; i++)
            slot[list[i].hash & mask] = i;
        for (int i = fixedWords; i < size; i++)
            list[i] = Word{ NULL, 0, i };
        next = fixedWords;
    }

    int size() const { return (int)list.size(); }

    // slot content for a word of `len` letters with hash h whose letters from the second on are `tail`; -1 if none
    int match(int s, u32 h, int len, const u8* tail) const
    {
        if (s < 0)
            return -1;
        const Word& w = list[s];
        if (w.hash != h || (w.lenIdx >> 24) != len)
            return -1;
        return memcmp(w.text + 1, tail, (size_t)(len - 1)) == 0 ? s : -2; // -2: same hash and length, other letters
    }

    // learn a word (its slot was empty): the oldest learnt word gives way
    void learn(const u8* text, u32 h, int len)
    {
        Word& w = list[next];
        if ((w.lenIdx & 0x7FFFF) >= fixedWords) {
            slot[w.hash & mask] = -1;
            w.text = text;
            w.hash = h;
            w.lenIdx = (len << 24) | next;
        }
        slot[h & mask] = next;
        next++;
        if (next >= size()) {
            if (size() >= MAX_WORDS) {
                next = fixedWords;
                return;
            }
            // the list doubles; every word re-enters the map in list order (later words win shared slots)
            const int old = size();
            list.resize((size_t)old * 2);
            for (int i = old; i < 2 * old; i++)
                list[i] = Word{ NULL, 0, i };
            for (int i = 0; i < old; i++)
                slot[list[i].hash & mask] = i;
        }
    }
};

// ---- statistics (TextCodec::computeStats) -----------------------------------------------------------------
u8 block_mode(const u8* p, int n, bool strict)
{
    if (!strict && n >= 4 && knz_magic_known(p))
        return MODE_NOT_TEXT;
    u32 f0[256];
    std::vector<u32> f1(65536, 0u);
    memset(f0, 0, sizeof(f0));
    u32 prev = 0;
    for (int i = 0; i < n; i++) {
        f0[p[i]]++;
        f1[(prev << 8) | p[i]]++;
        prev = p[i];
    }
    i64 letters = (i64)f0['\r'] + f0['\n'], ascii = 0;
    for (int i = 0; i < 128; i++) {
        if (is_letter((u8)i))
            letters += f0[i];
        ascii += f0[i];
    }
    const i64 binary = n - ascii;
    bool notText = binary > (n >> 2);
    if (!notText) {
        notText = letters < (n >> 2);
        if (strict)
            notText = notText || f0[0] >= (u32)(n / 100) || (ascii / 95) < (n / 100);
        else
            notText = notText || f0[' '] < (u32)(n / 50);
    }
    if (notText) {
        const int dt = knz_detect_simple_type(n, f0);
        if (dt != KDT_UNDEFINED)
            return (u8)(MODE_NOT_TEXT | dt);
        return knz_utf8_plausible(f0, f1.data(), n) ? (u8)(MODE_NOT_TEXT | KDT_UTF8) : MODE_NOT_TEXT;
    }
    u8 mode = 0;
    if (binary <= n - n / 10) { // mark-up: '<' and '>' about as frequent, and some '&a' '&g' '&l' '&q'
        const i64 lt = f0['<'], gt = f0['>'];
        const i64 amp = (i64)f1['&' * 256 + 'a'] + f1['&' * 256 + 'g'] + f1['&' * 256 + 'l'] + f1['&' * 256 + 'q'];
        i64 least = (n - binary) >> 9;
        if (least < 2)
            least = 2;
        if (lt >= least && gt >= least && amp > 0) {
            if (lt < gt) {
                if (lt >= gt - gt / 100)
                    mode |= MODE_XML;
            } else if (gt < lt) {
                if (gt >= lt - lt / 100)
                    mode |= MODE_XML;
            } else {
                mode |= MODE_XML;
            }
        }
    }
    if (f0['\r'] != 0 && f0['\r'] == f0['\n']) { // every CR followed by LF and every LF behind a CR
        bool paired = true;
        for (int i = 0; i < 256 && paired; i++) {
            if (i != '\n' && f1['\r' * 256 + i] != 0)
                paired = false;
            if (i != '\r' && f1[i * 256 + '\n'] != 0)
                paired = false;
        }
        if (paired)
            mode |= MODE_CRLF;
    }
    return mode;
}

// ---- the two wire variants -------------------------------------------------------------------------------
struct Variant1 {
    static const bool strict = true, escapes = true;
    static int log_slots(int blockSize)
    {
        if (blockSize < 8)
            return 13;
        const int lg = 31 - __builtin_clz((u32)(blockSize / 8));
        return lg > 26 ? 26 : (lg < 13 ? 13 : lg);
    }
    static const int headroom = 4;
    // index on 1..3 bytes: 7 bits, 7 + 7 bits, 5 + 7 + 7 bits
    static int put_index(u8* d, int v)
    {
        if (v < 128) {
            d[0] = (u8)v;
            return 1;
        }
        if (v < 128 * 128) {
            d[0] = (u8)(0x80 | (v >> 7)), d[1] = (u8)(v & 0x7F);
            return 2;
        }
        d[0] = (u8)(0xE0 | (v >> 14)), d[1] = (u8)(0x80 | (v >> 7)), d[2] = (u8)(v & 0x7F);
        return 3;
    }
    static int put_word(u8* d, int index, bool flipped)
    {
        d[0] = flipped ? TOKEN_FLIP : TOKEN_WORD;
        return 1 + put_index(d + 1, index);
    }
    // literal bytes between two word references; the two token values travel as references to the escape entries
    static int put_literals(const u8* s, int n, u8* d, int room, bool crlf, int fixedWords)
    {
        int k = 0;
        for (int i = 0; i < n; i++) {
            if (k >= room)
                return -1;
            const u8 c = s[i];
            if (c == TOKEN_WORD || c == TOKEN_FLIP) {
                d[k++] = TOKEN_WORD;
                const int idx = (c == TOKEN_WORD) ? fixedWords - 1 : fixedWords - 2;
                const int bytes = (idx >= 128) ? (idx >= 128 * 128 ? 3 : 2) : 1;
                if (k + bytes >= room)
                    return -1;
                k += put_index(d + k, idx);
            } else if (c == '\r') {
                if (!crlf)
                    d[k++] = c;
            } else {
                d[k++] = c;
            }
        }
        return k;
    }
};

struct Variant2 {
    static const bool strict = false, escapes = false;
    static int log_slots(int blockSize)
    {
        if (blockSize < 32)
            return 13;
        const int lg = 31 - __builtin_clz((u32)(blockSize / 32));
        return lg > 24 ? 24 : (lg < 13 ? 13 : lg);
    }
    static const int headroom = 3;
    // index + 1 behind a marker in the high bits: 10xxxxxx, 110xxxxx x, 1111xxxx x x (0x80 alone = case flip)
    static int put_word(u8* d, int index, bool flipped)
    {
        int k = 0;
        d[0] = 0x80;
        k += flipped ? 1 : 0;
        const int v = index + 1;
        if (v < 64) {
            d[k++] = (u8)(0x80 | v);
        } else if (v < 64 * 128) {
            d[k++] = (u8)(0xC0 | (v >> 8)), d[k++] = (u8)v;
        } else {
            d[k++] = (u8)(0xF0 | (v >> 16)), d[k++] = (u8)(v >> 8), d[k++] = (u8)v;
        }
        return k;
    }
    // bytes with the high bit set (and the escape itself) travel behind an escape byte
    static int put_literals(const u8* s, int n, u8* d, int room, bool crlf, int)
    {
        int k = 0;
        for (int i = 0; i < n; i++) {
            const u8 c = s[i];
            if (c == TOKEN_WORD) {
                if (k >= room - 1)
                    return -1;
                d[k++] = TOKEN_WORD, d[k++] = TOKEN_WORD;
            } else if (c == '\r') {
                if (!crlf) {
                    if (k >= room)
                        return -1;
                    d[k++] = c;
                }
            } else {
                if (c >= 128) {
                    if (k >= room)
                        return -1;
                    d[k++] = TOKEN_WORD;
                }
                if (k >= room)
                    return -1;
                d[k++] = c;
            }
        }
        return k;
    }
};

template <class V>
bool text_forward(const u8* src, int n, u8* dst, int cap, int* outLen, KnzPreCtx* pc)
{
    if (n < 1024 || cap < n)
        return false;
    if (pc->dataType != KDT_UNDEFINED && pc->dataType != KDT_TEXT && pc->dataType != KDT_BIN)
        return false; // binaries may still hold a good share of text
    const u8 mode = block_mode(src, n, V::strict);
    if (mode & MODE_NOT_TEXT) {
        pc->dataType = mode & 0x0F;
        return false;
    }
    pc->dataType = KDT_TEXT;
    Dictionary dict;
    dict.init(n, V::log_slots(pc->blockSize), V::escapes);
    const bool crlf = (mode & MODE_CRLF) != 0;
    const int room = n; // getMaxEncodedLength
    int s = 0, d = 0, pending = 0; // pending: first input byte not yet written
    dst[d++] = mode;
    while (s < n && src[s] == ' ') {
        dst[d++] = ' ';
        s++;
        pending++;
    }
    int delim = (s < n && is_letter(src[s])) ? s - 1 : s; // position of the previous delimiter
    u32 h1 = MULT1, h2 = MULT1;
    bool ok = true;
    for (; s < n; s++) {
        const u8 c = src[s];
        const int cls = char_class(c);
        if (cls == 0) {
            if (s - delim == 1) {
                h1 = hash_step(MULT1, c);
                h2 = hash_step(MULT1, (u8)(c ^ 0x20));
            } else {
                h1 = hash_step(h1, c);
                h2 = hash_step(h2, c);
            }
            continue;
        }
        const int len = s - delim - 1;
        if (cls > 0 && len >= 2 && len <= MAX_WORD) {
            const u8* word = src + delim + 1;
            const int s1 = dict.slot[h1 & dict.mask];
            int hit = dict.match(s1, h1, len, word + 1);
            if (hit == -1) // not as written: try with the first letter's case flipped
                hit = dict.match(dict.slot[h2 & dict.mask], h2, len, word + 1);
            const bool flipped = hit != s1; // (the reference compares the entries, not the hashes)
            if (hit < 0) {
                if ((len > 3 || (len == 3 && dict.next < 128 * 128)) && s1 < 0)
                    dict.learn(word, h1, len);
            } else {
                // literals up to the word; a single space between two references is implied
                if (pending != delim || src[delim] != ' ') {
                    const int k = V::put_literals(src + pending, delim + 1 - pending, dst + d, room - d, crlf, dict.fixedWords);
                    if (k < 0) {
                        ok = false;
                        break;
                    }
                    d += k;
                }
                if (d >= room - V::headroom) {
                    ok = false;
                    break;
                }
                d += V::put_word(dst + d, dict.list[hit].lenIdx & 0x7FFFF, flipped);
                pending = delim + 1 + len;
            }
        }
        delim = s;
    }
    if (ok) {
        const int k = V::put_literals(src + pending, n - pending, dst + d, room - d, crlf, dict.fixedWords);
        if (k < 0)
            ok = false;
        else
            d += k;
    }
    *outLen = d;
    return ok;
}

// Decode-side scan shared by both variants: learns the words the encoder learnt, from the literal text of the
// CODED block, at the positions where the encoder met them.
struct Learner {
    Dictionary dict;
    const u8* src;
    int delim;
    void before_delimiter(int s)
    {
        const int len = s - delim - 1;
        if (len < 3 || len > MAX_WORD)
            return;
        u32 h = MULT1;
        for (int i = delim + 1; i < s; i++)
            h = hash_step(h, src[i]);
        const int s1 = dict.slot[h & dict.mask];
        if (dict.match(s1, h, len, src + delim + 2) >= 0)
            return;
        if ((len > 3 || dict.next < 128 * 128) && s1 < 0)
            dict.learn(src + delim + 1, h, len);
    }
};

template <class V>
bool text_inverse(const u8* src, int n, u8* dst, int cap, int* outLen, int blockSize)
{
    if (n < 2)
        return false;
    Learner L;
    L.dict.init(cap, V::log_slots(blockSize), V::escapes);
    L.src = src;
    const bool crlf = (src[0] & MODE_CRLF) != 0;
    int s = 1, d = 0;
    L.delim = is_letter(src[s]) ? s - 1 : s;
    bool afterWord = false, ok = true;
    while (s < n && d < cap) {
        u8 c = src[s];
        const int cls = char_class(c);
        if (cls == 0) {
            dst[d++] = c;
            s++;
            continue;
        }
        if (cls > 0)
            L.before_delimiter(s);
        s++;
        int idx = -1;
        bool isRef = false;
        u8 flip = 0;
        if (V::escapes) { // variant 1: token byte, then the index
            if (c == TOKEN_WORD || c == TOKEN_FLIP) {
                isRef = true;
                if (s >= n) {
                    ok = false;
                    break;
                }
                idx = src[s++];
                if (idx >= 128) {
                    if (s >= n) {
                        ok = false;
                        break;
                    }
                    const int b2 = src[s++];
                    if (b2 >= 128) {
                        if (s >= n) {
                            ok = false;
                            break;
                        }
                        idx = ((idx & 0x1F) << 14) | ((b2 & 0x7F) << 7) | src[s++];
                    } else {
                        idx = ((idx & 0x7F) << 7) | b2;
                    }
                    if (idx >= L.dict.size()) {
                        ok = false;
                        break;
                    }
                }
                flip = (c == TOKEN_FLIP) ? 0x20 : 0;
            }
        } else if (c >= 0x80) { // variant 2: the index is in the byte itself
            isRef = true;
            if (c == 0x80) {
                flip = 0x20;
                if (s >= n) {
                    ok = false;
                    break;
                }
                c = src[s++];
            }
            idx = c & 0x7F;
            if (idx >= 64) {
                if (idx >= 112) {
                    if (s + 1 >= n) {
                        ok = false;
                        break;
                    }
                    idx = ((idx & 0x0F) << 16) | (src[s] << 8) | src[s + 1];
                    s += 2;
                } else {
                    if (s >= n) {
                        ok = false;
                        break;
                    }
                    idx = ((idx & 0x1F) << 8) | src[s++];
                }
                if (idx > L.dict.size()) {
                    ok = false;
                    break;
                }
            } else if (idx == 0) {
                ok = false;
                break;
            }
            idx--;
        }
        if (isRef) {
            if (idx < 0 || idx >= L.dict.size()) {
                ok = false;
                break;
            }
            const Word& w = L.dict.list[idx];
            const int len = (w.lenIdx >> 24) & 0xFF;
            if (len > 1) {
                if (afterWord)
                    dst[d++] = ' '; // the implied space between two references
                afterWord = true;
                L.delim = s;
            } else {
                if (len == 0) {
                    ok = false;
                    break;
                }
                afterWord = false; // an escaped token byte
                L.delim = s - 1;
            }
            if (d + len > cap) {
                ok = false;
                break;
            }
            memcpy(dst + d, w.text, (size_t)len);
            dst[d] ^= flip;
            d += len;
            continue;
        }
        if (!V::escapes && c == TOKEN_WORD) { // variant 2: the next byte is a literal
            if (s >= n) {
                ok = false;
                break;
            }
            dst[d++] = src[s++];
        } else {
            if (crlf && c == '\n') {
                dst[d++] = '\r';
                if (d >= cap) {
                    ok = false;
                    break;
                }
            }
            dst[d++] = c;
        }
        afterWord = false;
        L.delim = s - 1;
    }
    *outLen = d;
    return ok && s == n;
}

inline bool variant2_for(int eType)
{
    // NONE, HUFFMAN, ANS0 (and RANGE, id 3, which this library does not code) select variant 2
    return eType == E_RAW || eType == E_HUF || eType == E_ANS0 || eType == 3;
}

} // namespace

bool knz_text_forward(const u8* src, int n, u8* dst, int cap, int* outLen, KnzPreCtx* pc)
{
    return variant2_for(pc->eType) ? text_forward<Variant2>(src, n, dst, cap, outLen, pc)
                                   : text_forward<Variant1>(src, n, dst, cap, outLen, pc);
}

bool knz_text_inverse(const u8* src, int n, u8* dst, int cap, int* outLen, int blockSize, int eType)
{
    return variant2_for(eType) ? text_inverse<Variant2>(src, n, dst, cap, outLen, blockSize)
                               : text_inverse<Variant1>(src, n, dst, cap, outLen, blockSize);
}
