// TEST SCAFFOLDING: compiles the device decoder source (inflate_core.cuh) in
// PP_HOST_EMU mode — lanes run one after another — so the DEFLATE logic can be
// checked against zlib on a machine without a GPU.  Never linked into the product.
#define PP_HOST_EMU 1
#include <stdint.h>
#include <stdlib.h>
struct uint4 { uint32_t x, y, z, w; };
#include "../../parallelparsing_b200/csrc/inflate_core.cuh"

extern "C" {
// comp: compressed buffer padded to a multiple of the tile size (+1 tile); slot:
// [lead_len window bytes][out_len output][128 pad].  Returns status; fills res[4]:
// produced, newlines, min_byte, end_bit (low 32).
int emu_inflate_chunk(const uint8_t *comp, uint64_t comp_bytes, uint64_t in_bit, uint64_t in_limit,
                      uint8_t *slot, const uint8_t *lead, uint32_t lead_len, uint32_t out_len, uint64_t *res)
{
    ppinf::ChunkDesc d;
    d.in_bit = in_bit; d.in_limit = in_limit; d.slot_off = 0; d.lead_src = 0;
    d.lead_len = lead_len; d.out_len = out_len; d.prefix_len = 0; d.prefix_nl = 0;
    ppinf::ChunkResult r;
    ppinf::inflate_chunk(d, comp, comp_bytes, slot, lead, r);
    res[0] = r.produced; res[1] = r.newlines; res[2] = r.min_byte; res[3] = r.end_bit;
    return r.status;
}
int emu_tile_bytes(void) { return ppinf::kTileBytes; }
}
