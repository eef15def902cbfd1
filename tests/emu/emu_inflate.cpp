// TEST SCAFFOLDING: compiles the device decoder source (inflate_core.cuh) in
// PP_HOST_EMU mode — the CTA's threads run one after another, phase by phase — so
// the DEFLATE logic can be checked against zlib on a machine without a GPU.
// Never linked into the product.
#define PP_HOST_EMU 1
#include <stdint.h>
#include <stdlib.h>
struct uint4 { uint32_t x, y, z, w; };
#include "../../parallelparsing_b200/csrc/inflate_core.cuh"


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
int emu_subw(void) { return ppinf::kSubW; }
void emu_stats(uint64_t *out, int reset)
{
    for (int i = 0; i < 8; i++) { out[i] = ppinf::g_stat[i]; if (reset) ppinf::g_stat[i] = 0; }
}
}
