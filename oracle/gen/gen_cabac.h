/* gen_cabac.h -- TEST INFRASTRUCTURE ONLY (bitstream fabrication; never linked into the product).
 *
 * Force-included (gcc -include) in front of the reference's own vcl_*.c when they are compiled, in place, into
 * oracle/_ref/libovvc_gen.so.  It takes the place of libovvc/cabac_internal.h (same include guard), so that every context-coded
 * and bypass bin the reference's PARSER asks for (ovcabac_ae_read / ovcabac_bypass_read, cabac_internal.h:45-161) is not decoded
 * from a stream but CHOSEN by oracle/gen/gen_cabac.c, which at the same time arithmetic-ENCODES the chosen bin with the
 * context state the parser holds.  The parser itself therefore drives the syntax (which element comes next, which context it
 * uses, when a CTU ends); what comes out is a slice-data payload that the unmodified decoder parses back into exactly the
 * same bins.  No reference source is copied: the parser is compiled from /root/reference where it lies. */
#ifndef CABAC_INTERNAL_H
#define CABAC_INTERNAL_H
#include "ovutils.h"
#include "vcl_cabac.h"

#define NB_CABAC_BITS 16
#define CABAC_MASK ((1 << NB_CABAC_BITS) - 1)

extern const uint8_t lps_renorm_table[64];
extern const uint8_t range_lps_lut[512];

uint8_t ovgen_ae_read(OVCABACCtx *cabac_ctx, uint64_t *cabac_state);
uint8_t ovgen_bypass_read(OVCABACCtx *cabac_ctx);

static inline uint8_t ovcabac_ae_read(OVCABACCtx *const cabac_ctx, uint64_t *const cabac_state)
{
    return ovgen_ae_read(cabac_ctx, cabac_state);
}
static inline uint8_t ovcabac_bypass_read(OVCABACCtx *const cabac_ctx)
{
    return ovgen_bypass_read(cabac_ctx);
}
#endif
