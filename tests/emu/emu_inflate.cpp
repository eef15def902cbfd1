// TEST SCAFFOLDING: compiles the device decoder source (inflate_core.cuh) in
// PP_HOST_EMU mode — the CTA's threads run one after another, phase by phase — so
// the DEFLATE logic can be checked against zlib on a machine without a GPU.
// Never linked into the product.
#define PP_HOST_EMU 1
#include <stdint.h>
#include <stdlib.h>
struct uint4 { uint32_t x, y, z, w; };
#include "../../parallelparsing_b200/csrc/inflate_core.cuh"
#include "../../parallelparsing_b200/csrc/blockscan_core.cuh"


extern "C" {
// comp: compressed buffer (comp_bytes a multiple of 16); slot: [lead_len window bytes]
// [out_len output][128 pad].  T = emulated CTA size.  Returns status; fills res[4]:
// produced, newlines, min_byte, end_bit.
int emu_inflate_chunk(int T, const uint8_t *comp, uint64_t comp_bytes, uint64_t in_bit, uint64_t in_limit,
                      uint8_t *slot, const uint8_t *lead, uint32_t lead_len, uint32_t out_len, uint64_t *res)
{
    ppinf::g_T = T;
    ppinf::ChunkDesc d;
    d.in_bit = in_bit; d.in_limit = in_limit; d.slot_off = 0; d.lead_src = 0;
    d.lead_len = lead_len; d.out_len = out_len; d.prefix_len = 0; d.prefix_nl = 0;
    ppinf::ChunkResult r;
    uint8_t *raw = (uint8_t *)aligned_alloc(128, ppinf::sm_bytes_for(T));
    uint32_t *map = (uint32_t *)aligned_alloc(128, (ppinf::scratch_words_for(T) * 4 + 127) / 128 * 128);
    memset(map, 0xff, ppinf::scratch_words_for(T) * 4);  // scratch is never zero on the device either
    memset(raw, 0xff, ppinf::sm_bytes_for(T));
    ppinf::Sm sm;
    ppinf::sm_carve(sm, raw, T);
    uint32_t phase = 0;
    ppinf::inflate_chunk(sm, d, comp, comp_bytes, slot, lead, map, r, phase);
    free(raw);
    free(map);
    res[0] = r.produced; res[1] = r.newlines; res[2] = r.min_byte; res[3] = r.end_bit;
    return r.status;
}
// Pull mode: a window re-uses the staged bytes the window before it left behind RESOLVE's scratch.
int emu_inflate_chunk_pull(int T, const uint8_t *comp, uint64_t comp_bytes, uint64_t in_bit, uint64_t in_limit,
                           uint8_t *slot, const uint8_t *lead, uint32_t lead_len, uint32_t out_len, uint64_t *res)
{
    ppinf::g_T = T;
    ppinf::ChunkDesc d;
    d.in_bit = in_bit; d.in_limit = in_limit; d.slot_off = 0; d.lead_src = 0;
    d.lead_len = lead_len; d.out_len = out_len; d.prefix_len = 0; d.prefix_nl = 0;
    ppinf::ChunkResult r;
    uint8_t *raw = (uint8_t *)aligned_alloc(128, ppinf::sm_bytes_for(T));
    uint32_t *map = (uint32_t *)aligned_alloc(128, (ppinf::scratch_words_for(T) * 4 + 127) / 128 * 128);
    memset(map, 0xff, ppinf::scratch_words_for(T) * 4);
    memset(raw, 0xff, ppinf::sm_bytes_for(T));
    ppinf::Sm sm;
    ppinf::sm_carve(sm, raw, T);
    uint32_t phase = 0;
    ppinf::inflate_chunk<false, true>(sm, d, comp, comp_bytes, slot, lead, map, r, phase);
    free(raw);
    free(map);
    res[0] = r.produced; res[1] = r.newlines; res[2] = r.min_byte; res[3] = r.end_bit;
    return r.status;
}
// Dual output (GPU CreateIndex): one decode, resolved against two histories; slot holds both outputs,
// the second one slot_delta bytes after the first; lead holds both histories, lead_len apart.
int emu_inflate_chunk_dual(int T, const uint8_t *comp, uint64_t comp_bytes, uint64_t in_bit, uint64_t in_limit,
                           uint8_t *slot, uint64_t slot_delta, const uint8_t *lead, uint32_t lead_len, uint32_t out_len,
                           uint64_t *res)
{
    ppinf::g_T = T;
    ppinf::ChunkDesc d;
    d.in_bit = in_bit; d.in_limit = in_limit; d.slot_off = 0; d.lead_src = 0;
    d.lead_len = lead_len; d.out_len = out_len; d.prefix_len = 0; d.prefix_nl = 0;
    ppinf::ChunkResult r;
    uint8_t *raw = (uint8_t *)aligned_alloc(128, ppinf::sm_bytes_for(T));
    uint32_t *map = (uint32_t *)aligned_alloc(128, (ppinf::scratch_words_for(T) * 4 + 127) / 128 * 128);
    memset(map, 0xff, ppinf::scratch_words_for(T) * 4);
    memset(raw, 0xff, ppinf::sm_bytes_for(T));
    ppinf::Sm sm;
    ppinf::sm_carve(sm, raw, T);
    uint32_t phase = 0;
    ppinf::inflate_chunk<true>(sm, d, comp, comp_bytes, slot, lead, map, r, phase, nullptr, ppinf::DualOut{slot_delta, lead_len});
    free(raw);
    free(map);
    res[0] = r.produced; res[1] = r.newlines; res[2] = r.min_byte; res[3] = r.end_bit;
    return r.status;
}
// Block scanner (GPU-assisted CreateIndex, first slice): one segment, emulated CTA of T threads.
// recs: rec_cap x {bit, out}; res[5]: first_bit, land_bit, out_bytes, nrec, status.
int emu_scan_segment(int T, const uint8_t *comp, uint64_t comp_bytes, uint64_t start_bit, uint64_t end_bit, int search,
                     uint64_t *recs, uint32_t rec_cap, uint64_t *res)
{
    ppinf::g_T = T;
    uint8_t *raw = (uint8_t *)aligned_alloc(128, ppinf::sm_bytes_for(T));
    memset(raw, 0xff, ppinf::sm_bytes_for(T));
    ppinf::Sm sm;
    ppinf::sm_carve(sm, raw, T);
    ppinf::ScanSegIn in;
    in.start_bit = start_bit; in.end_bit = end_bit; in.search = (uint32_t)search; in.rec_off = 0; in.rec_cap = rec_cap; in.pad = 0;
    ppinf::ScanSegOut out;
    uint32_t phase = 0;
    ppinf::scan_segment(sm, in, comp, comp_bytes, 0, (ppinf::BlockRec *)recs, out, phase);
    free(raw);
    res[0] = out.first_bit; res[1] = out.land_bit; res[2] = out.out_bytes; res[3] = out.nrec; res[4] = (uint64_t)(int64_t)out.status;
    return out.status;
}
int emu_probe_dynamic_header(const uint8_t *comp, uint64_t comp_bytes, uint64_t bit)
{
    ppinf::BitPeek bp;
    bp.w = (const uint32_t *)comp; bp.nw = comp_bytes / 4; bp.shift = 0;
    return ppinf::probe_dynamic_header(bp, bit) ? 1 : 0;
}
int emu_subw(void) { return ppinf::kSubW; }
void emu_stats(uint64_t *out, int reset)
{
    for (int i = 0; i < 8; i++) { out[i] = ppinf::g_stat[i]; if (reset) ppinf::g_stat[i] = 0; }
}
}
