// GPU-assisted CreateIndex, first slice: where do the deflate blocks of a gzip member start, and how
// many bytes does each produce?  (Core.BuildDeflateIndex, Decompressor/Core.cs:14-131, gets exactly
// this from zlib's inflate(Z_BLOCK) — a serial pass over the whole file at ~350 MB/s per host core;
// its checkpoints may only sit at block starts, Core.cs:98.)
//
// Finding block boundaries needs the Huffman decode but NOT the LZ77 history: a block's header and
// symbols can be walked from its first bit alone.  So the compressed stream is cut into SEGMENTS, one
// CTA each:
//   SEARCH  (all segments but the first, whose start is known: the end of the gzip header) every
//           thread probes one bit position for "a valid dynamic-Huffman block header starts here" —
//           block type, symbol counts, a complete code-length code, well-formed run lengths, an
//           end-of-block code, complete (or single-code) literal/length and distance sets: zlib's own
//           rules.  The first position that passes is taken as a block start (pugz / rapidgzip style);
//   WALK    from there the CTA walks block after block — header, tables, and per window the
//           GUESS / SYNC / SCAN half of the inflate kernel (count_window: no tokens, no output) —
//           recording (first bit, bytes produced so far) per block, until a block starts at or past the
//           segment's end: that position is where the NEXT segment's first block must be.
// The host stitches the segments: a segment is accepted when the walk before it LANDS exactly on
// the block start its search found; anything else (a stored/fixed block at the seam, which the search
// cannot see, or a false positive) is re-walked from the landing position without a search.
#pragma once
#include "inflate_core.cuh"

namespace ppinf {

struct ScanSegIn {
    uint64_t start_bit;   // first bit of the segment (search) / of its first block (no search)
    uint64_t end_bit;     // the walk stops at the first block start >= end_bit
    uint32_t search;      // 1: probe for the first dynamic block header in [start_bit, end_bit)
    uint32_t rec_off;     // where this segment's records go in the record array
    uint32_t rec_cap;
    uint32_t pad;
};
struct ScanSegOut {
    uint64_t first_bit;   // where the walk started (~0: the search found nothing)
    uint64_t land_bit;    // first block start >= end_bit, or the bit after the final block
    uint64_t out_bytes;   // bytes produced by the blocks walked
    uint32_t nrec;        // records written (block starts in [first_bit, land_bit))
    int32_t status;       // 0, 1 = the stream's final block was walked, -3 = invalid data, -5 = record array full
};
struct BlockRec {
    uint64_t bit;         // first bit of the block header (the BFINAL bit)
    uint64_t out;         // bytes produced by the segment's blocks before this one
};

// ---- SEARCH probe: one thread, bits straight from global memory -------------------------------
struct BitPeek {
    const uint32_t *w;    // 4-byte aligned compressed words
    uint64_t nw;          // readable words; past them the stream reads as zeros
    uint64_t shift;       // bits to add to every position (alignment of the base pointer)
};
PP_HD uint32_t bp_word(const BitPeek &b, uint64_t i) { return i < b.nw ? b.w[i] : 0u; }
PP_HD uint32_t bp_peek(const BitPeek &b, uint64_t pos, uint32_t n)  // n <= 25
{
    pos += b.shift;
    const uint64_t i = pos >> 5;
    const uint32_t s = (uint32_t)pos & 31u;
    const uint64_t v = (uint64_t)bp_word(b, i) | ((uint64_t)bp_word(b, i + 1) << 32);
    return (uint32_t)(v >> s) & ((1u << n) - 1u);
}

// The cheap front of the probe: block type, symbol counts, and a COMPLETE code-length code (zlib
// inflate_table, type CODES).  About 1 % of random bit positions pass.
PP_HD bool probe_cheap(const BitPeek &b, uint64_t pos)
{
    const uint32_t h = bp_peek(b, pos, 17);
    if (((h >> 1) & 3u) != 2u) return false;                         // BTYPE
    const uint32_t nlen = ((h >> 3) & 31u) + 257u, ndist = ((h >> 8) & 31u) + 1u, ncode = ((h >> 13) & 15u) + 4u;
    if (nlen > 286u || ndist > 30u) return false;
    uint64_t p = pos + 17u;
    int32_t left = 1 << 7;                                           // Kraft sum in units of 2^-7
    for (uint32_t i = 0; i < ncode; i += 8u) {                       // eight 3-bit lengths per peek
        const uint32_t m = ncode - i < 8u ? ncode - i : 8u;
        const uint32_t v = bp_peek(b, p, 3u * m);
        p += 3u * m;
        for (uint32_t k = 0; k < m; k++) {
            const uint32_t l = (v >> (3u * k)) & 7u;
            if (l) left -= 1 << (7 - l);
        }
    }
    return left == 0;
}

// Is `pos` the first bit of a VALID dynamic-Huffman block header (RFC 1951 3.2.7 under zlib's checks)?
PP_HD bool probe_dynamic_header(const BitPeek &b, uint64_t pos)
{
    const uint32_t h = bp_peek(b, pos, 17);
    if (((h >> 1) & 3u) != 2u) return false;                         // BTYPE
    const uint32_t nlen = ((h >> 3) & 31u) + 257u, ndist = ((h >> 8) & 31u) + 1u, ncode = ((h >> 13) & 15u) + 4u;
    if (nlen > 286u || ndist > 30u) return false;                    // zlib: too many length or distance symbols
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t cl[19];
    for (int i = 0; i < 19; i++) cl[i] = 0;
    uint64_t p = pos + 17u;
    uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t i = 0; i < ncode; i++) {
        const uint32_t l = bp_peek(b, p, 3);
        p += 3u;
        cl[order[i]] = (uint8_t)l;
        cnt[l]++;
    }
    {   // the code-length code must be complete (zlib inflate_table, type CODES)
        int left = 1;
        for (int l = 1; l <= 7; l++) {
            left = (left << 1) - (int)cnt[l];
            if (left < 0) return false;
        }
        if (left != 0) return false;
    }
    // canonical decode of the code-length code: first code and first sorted index per length
    uint8_t sorted[19];
    uint32_t offs[9];
    offs[1] = 0;
    for (int l = 1; l <= 7; l++) offs[l + 1] = offs[l] + cnt[l];
    {
        uint32_t fill[8];
        for (int l = 0; l < 8; l++) fill[l] = offs[l < 1 ? 1 : l];
        for (int s = 0; s < 19; s++)
            if (cl[s]) sorted[fill[cl[s]]++] = (uint8_t)s;
    }
    uint32_t lcnt[16], dcnt[16];
    for (int l = 0; l < 16; l++) lcnt[l] = dcnt[l] = 0;
    const uint32_t total = nlen + ndist;
    uint32_t have = 0, prev = 0, lensyms = 0;
    bool eob = false;
    // running Kraft sums (in units of 2^-15): an over-subscribed set is rejected as soon as it overflows —
    // at a wrong position the lengths are noise and that happens within a few dozen symbols, not after 300
    int32_t room_l = 1 << 15, room_d = 1 << 15;
    while (have < total) {
        // one symbol of the code-length code, bit by bit (codes are MSB first, the stream LSB first)
        const uint32_t bits = bp_peek(b, p, 7);
        uint32_t code = 0, first = 0, idx = 0, sym = 99, used = 0;
        for (uint32_t l = 1; l <= 7; l++) {
            code |= (bits >> (l - 1u)) & 1u;
            const uint32_t c = cnt[l];
            if (code - first < c) { sym = sorted[idx + (code - first)]; used = l; break; }
            idx += c;
            first = (first + c) << 1;
            code <<= 1;
        }
        if (sym == 99u) return false;                                // (cannot happen for a complete code)
        p += used;
        uint32_t rep = 1, val = sym;
        if (sym == 16u) {
            if (have == 0u) return false;                            // repeat with no previous length
            rep = 3u + bp_peek(b, p, 2); p += 2u; val = prev;
        } else if (sym == 17u) {
            rep = 3u + bp_peek(b, p, 3); p += 3u; val = 0;
        } else if (sym == 18u) {
            rep = 11u + bp_peek(b, p, 7); p += 7u; val = 0;
        }
        if (have + rep > total) return false;                        // repeat past the end
        if (sym < 16u) prev = sym;
        else if (sym != 16u) prev = 0;
        for (uint32_t r = 0; r < rep; r++) {
            const uint32_t at = have + r;
            if (at < nlen) {
                lcnt[val]++;
                if (at == 256u && val) eob = true;
                if (at > 256u && val) lensyms++;
                if (val) room_l -= 1 << (15 - val);
            } else {
                dcnt[val]++;
                if (val) room_d -= 1 << (15 - val);
            }
        }
        if (room_l < 0 || room_d < 0) return false;                  // over-subscribed
        have += rep;
    }
    if (!eob) return false;                                          // missing end-of-block code
    for (int set = 0; set < 2; set++) {                              // literal/length, distance
        const uint32_t *c = set ? dcnt : lcnt;
        int left = 1, maxlen = 0;
        for (int l = 1; l <= 15; l++) {
            if (c[l]) maxlen = l;
            left = (left << 1) - (int)c[l];
            if (left < 0) return false;                              // over-subscribed
        }
        if (left > 0 && maxlen != 1 && !(set == 1 && maxlen == 0)) return false;  // incomplete
        // Stricter than zlib, which is fine for a SEARCH (a true start that is turned down only costs a re-walk
        // of its seam): length symbols with a code but not one distance code — zlib would accept the header and
        // fail at the first match; no compressor writes it, noise does (seen on the 10 M-read corpus).
        if (set == 1 && maxlen == 0 && lensyms) return false;
    }
    return true;
}

PP_DEV uint32_t atomic_inc_u32(uint32_t *p)   // returns the old value
{
#ifdef PP_HOST_EMU
    return (*p)++;
#else
    return atomicAdd(p, 1u);
#endif
}

// ---- WALK: the CTA follows the blocks of one segment -------------------------------------------
PP_DEV void scan_segment(const Sm &sm, const ScanSegIn &in, const uint8_t *comp, uint64_t comp_bytes, uint64_t shift_bits,
                         BlockRec *recs, ScanSegOut &out, uint32_t &stage_phase)
{
    const int T = PP_NT;
    uint64_t bit = in.start_bit;
    // SEARCH
    if (in.search) {
        BitPeek bp;
        bp.w = reinterpret_cast<const uint32_t *>(comp);
        bp.nw = comp_bytes / 4u;
        bp.shift = shift_bits;
        uint64_t found = ~0ull;
        const bool allow_final = (in.pad & 1u) != 0u;
        // Rounds of T x kSpan positions.  Pass 1: every thread runs the cheap filter over kSpan positions
        // and appends the survivors (~1 %) to a list in shared memory; pass 2: one survivor per THREAD
        // through the full probe (a serial decode of up to ~300 code lengths) — all lanes busy with
        // their own candidate instead of one lane per warp grinding while the others wait.
        // kSpan 128: a block start is ~half a block away (a few hundred thousand positions), so a round of
        // T x 128 positions costs little extra search and quarters the barriers and full-probe tails.
        constexpr uint32_t kSpan = 128, kCap = 4096;                 // list: kCap u32 in the resolve tile buffer (>= 16 KB at T >= 512)
        uint32_t *cand = reinterpret_cast<uint32_t *>(sm.res);
        const uint32_t cap = (uint32_t)T * kTileB / 2u < kCap ? (uint32_t)T * kTileB / 2u : kCap;
        for (uint64_t base = in.start_bit; base < in.end_bit && found == ~0ull; base += (uint64_t)T * kSpan) {
            PP_T0_BEGIN
            sm.u[8] = 0xffffffffu;   // smallest valid position of the round (relative to base)
            sm.u[9] = 0;             // survivors
            PP_T0_END
            PP_SYNC();
            PP_FOR_T(t)
            for (uint32_t i = 0; i < kSpan; i++) {
                const uint32_t rel = i * (uint32_t)T + (uint32_t)t;
                const uint64_t p = base + rel;
                // (BFINAL set: only the stream's last block may say so, so only searches near the stream's end take
                // such a candidate — in.pad bit 0; half of the false candidates go with it)
                if (p < in.end_bit && (allow_final || bp_peek(bp, p, 1) == 0u) && probe_cheap(bp, p)) {
                    const uint32_t at = atomic_inc_u32(&sm.u[9]);
                    if (at < cap) cand[at] = rel;
                }
            }
            PP_END_T
            PP_SYNC();
            const uint32_t nc = sm.u[9];
            if (nc <= cap) {
                PP_FOR_T(t)
                for (uint32_t i = (uint32_t)t; i < nc; i += (uint32_t)T) {
                    const uint32_t rel = cand[i];
                    if (probe_dynamic_header(bp, base + rel)) PP_ATOMIC_MIN(&sm.u[8], rel);
                }
                PP_END_T
            } else {   // the list overflowed (adversarial input): every position through the full probe
                PP_FOR_T(t)
                for (uint32_t i = 0; i < kSpan; i++) {
                    const uint32_t rel = i * (uint32_t)T + (uint32_t)t;
                    if (base + rel < in.end_bit && (allow_final || bp_peek(bp, base + rel, 1) == 0u) && probe_dynamic_header(bp, base + rel))
                        PP_ATOMIC_MIN(&sm.u[8], rel);
                }
                PP_END_T
            }
            PP_SYNC();
            if (sm.u[8] != 0xffffffffu) found = base + sm.u[8];
            PP_SYNC();
        }
        if (found == ~0ull) {
            PP_T0_BEGIN
            out.first_bit = ~0ull; out.land_bit = ~0ull; out.out_bytes = 0; out.nrec = 0; out.status = 0;
            PP_T0_END
            return;
        }
        bit = found;
    }
    const uint64_t first_bit = bit;
    const uint32_t cww = cw_words_for(T);
    uint64_t produced = 0;
    uint32_t nrec = 0;
    int status = 0;
    bool need_header = true, last = false;
    for (;;) {
        if (need_header && bit >= in.end_bit) break;                       // landed
        if (((bit + shift_bits) >> 3) >= comp_bytes) { status = -3; break; } // ran off the end of the file
        const uint64_t base_byte = ((bit + shift_bits) >> 3) & ~(uint64_t)15;
        if (!stage_window(sm, comp, comp_bytes, base_byte, cww, stage_phase)) { status = -100; break; }
        uint32_t s0 = (uint32_t)(bit + shift_bits - base_byte * 8u);
        if (need_header) {
            if (nrec >= in.rec_cap) { status = -5; break; }
            PP_T0_BEGIN
            recs[in.rec_off + nrec].bit = bit;
            recs[in.rec_off + nrec].out = produced;
            PP_T0_END
            nrec++;
            const uint32_t hdr = peek_bits(sm.cw, s0, 3);
            s0 += 3;
            last = (hdr & 1u) != 0;
            const uint32_t type = hdr >> 1;
            if (type == 0u) {
                const uint32_t bpos = (s0 + 7u) & ~7u;
                const uint32_t len = peek_bits(sm.cw, bpos, 16), nlen = peek_bits(sm.cw, bpos + 16u, 16);
                if ((len ^ 0xffffu) != nlen) { status = -3; break; }
                const uint64_t byte0 = base_byte + (bpos >> 3) + 4u;
                if (byte0 + len > comp_bytes) { status = -3; break; }
                produced += len;
                bit = (byte0 + len) * 8u - shift_bits;
                if (last) { status = 1; break; }
                continue;
            }
            int rc;
            if (type == 1u) rc = fixed_tables(sm);
            else if (type == 2u) rc = dynamic_tables(sm, s0, &s0);
            else rc = -3;
            if (rc) { status = rc; break; }
            need_header = false;
        }
        const WindowCount w = count_window(sm, s0, 0xffffffffu);
        produced += w.produced;
        bit = base_byte * 8u + w.next_bit - shift_bits;
        if (w.flag == F_BAD) { status = -3; break; }
        if (w.flag == F_EOB) {
            need_header = true;
            if (last) { status = 1; break; }
        }
    }
    PP_SYNC();
    PP_T0_BEGIN
    out.first_bit = first_bit;
    out.land_bit = bit;
    out.out_bytes = produced;
    out.nrec = nrec;
    out.status = status;
    PP_T0_END
    PP_SYNC();
}

}  // namespace ppinf
