/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's scan algorithm for sizes the numpy oracle
 * (oracle/scan_oracle.py) cannot reach.  Never linked into the product library.
 *
 * Restates, per row and PWM (smeagol/models.py:50-68 conv as a window sum, :108 strict '>' against fp64
 * thresholds, :110 end trim; smeagol/encode.py:11-29 soft one-hot):
 *   score32 = sum_k mix32(code[p+k], T[k])   ascending k on the FORWARD string ("kernel" order; for the '-' strand
 *             with Trc[k][b] = T[W-1-k][3-b] and the position reported in reverse-complement coordinates L-p-W)
 *   score64 = the same sum in double from the fp32 operands
 *   hit     = |score64 - thr| <= band ? score64 > thr : (double)score32 > thr         (fp64 adjudication)
 * mix32(c, t) = ((e0*t0 + e1*t1) + e2*t2) + e3*t3, every product and sum rounded (compile with -ffp-contract=off).
 * Build: make -C oracle   ->  oracle/_build/liboracle.so */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define T3 0.3333333432674407958984375f
static const float EMB[17][4] = {
    {0, 0, 0, 0}, {1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}, {0, 0, 0, 1}, {.25f, .25f, .25f, .25f},
    {.5f, 0, 0, .5f}, {0, .5f, .5f, 0}, {.5f, .5f, 0, 0}, {0, 0, .5f, .5f}, {.5f, 0, .5f, 0}, {0, .5f, 0, .5f},
    {0, T3, T3, T3}, {T3, 0, T3, T3}, {T3, T3, 0, T3}, {T3, T3, T3, 0}};

/* Scan n_rows forward rows (integer codes 0..16, row r = codes[row_off[r] .. +row_len[r])) with n_motifs PWMs
 * (tab: concatenated (W,4) fp32 tables, tab_off[m] = first table row of PWM m).  strands: bit0 forward, bit1 rc.
 * out_rows[r*2+st] = output row index used in the hit key
 *   key = out_row << (pos_bits+motif_bits) | pos << motif_bits | motif.
 * Outputs: counts[r][m][st] (int64), totals[0]=hits, [1]=near-threshold sites, [2]=sum of keys mod 2^64,
 * [3]=xor of keys, [4]=near sites whose fp32 and fp64 decisions differ.
 * If hit_cap > 0 the hits (key, score32) are also stored, UNSORTED, up to hit_cap; returns 0, or 1 on overflow. */
int orc_scan(const uint8_t* codes, const int64_t* row_off, const int64_t* row_len, int n_rows, const float* tab,
             const int32_t* widths, const int64_t* tab_off, int n_motifs, const double* thr, const double* band,
             int strands, const int32_t* out_rows, int pos_bits, int motif_bits, int64_t* counts, uint64_t* totals,
             uint64_t* hit_keys, float* hit_scores, int64_t hit_cap)
{
    uint64_t n_hits = 0, n_near = 0, ksum = 0, kxor = 0, n_flip = 0;
    int overflow = 0;
    memset(counts, 0, sizeof(int64_t) * (size_t)n_rows * n_motifs * 2);
#pragma omp parallel for collapse(2) schedule(dynamic, 4) reduction(+ : n_hits, n_near, ksum, n_flip) reduction(^ : kxor)
    for (int r = 0; r < n_rows; ++r) {
        for (int m = 0; m < n_motifs; ++m) {
            const uint8_t* c = codes + row_off[r];
            const int64_t L = row_len[r];
            const int W = widths[m];
            const float* T = tab + tab_off[m] * 4;
            int finite_tab = 1;
            for (int i = 0; i < 4 * W; ++i) finite_tab &= isfinite(T[i]) ? 1 : 0;
            for (int st = 0; st < 2; ++st) {
                if (!((strands >> st) & 1)) continue;
                int64_t cnt = 0;
                for (int64_t p = 0; p + W <= L; ++p) {
                    float s32 = 0.f;
                    double s64 = 0.0;
                    for (int k = 0; k < W; ++k) {
                        const float* tr = T + (st ? (W - 1 - k) : k) * 4;
                        const float t0 = st ? tr[3] : tr[0], t1 = st ? tr[2] : tr[1], t2 = st ? tr[1] : tr[2],
                                    t3 = st ? tr[0] : tr[3];
                        const int code = c[p + k];
                        float term;
                        double term64;
                        if (code >= 1 && code <= 5 && finite_tab) {
                            /* plain base: the mix reduces to the selected weight exactly (1*t + 0 + 0 + 0, finite t) */
                            const int b = code == 5 ? 3 : code - 1;
                            term = b == 0 ? t0 : b == 1 ? t1 : b == 2 ? t2 : t3;
                            term64 = (double)term;
                        } else {
                            const float* e = EMB[code];
                            term = ((e[0] * t0 + e[1] * t1) + e[2] * t2) + e[3] * t3;
                            term64 = (((double)e[0] * (double)t0 + (double)e[1] * (double)t1) +
                                      (double)e[2] * (double)t2) + (double)e[3] * (double)t3;
                        }
                        s32 = k ? s32 + term : term;
                        s64 = k ? s64 + term64 : term64;
                    }
                    const int near = fabs(s64 - thr[m]) <= band[m];
                    const int hit32 = (double)s32 > thr[m], hit64 = s64 > thr[m];
                    if (near) { ++n_near; if (hit32 != hit64) ++n_flip; }
                    if (near ? hit64 : hit32) {
                        const int64_t pos = st ? (L - p - W) : p;
                        const uint64_t key = ((uint64_t)(uint32_t)out_rows[r * 2 + st] << (pos_bits + motif_bits)) |
                                             ((uint64_t)pos << motif_bits) | (uint64_t)m;
                        ++cnt; ++n_hits; ksum += key; kxor ^= key;
                        if (hit_cap > 0) {
                            int64_t slot;
#pragma omp atomic capture
                            slot = totals[5]++;
                            if (slot < hit_cap) { hit_keys[slot] = key; hit_scores[slot] = s32; }
                            else overflow = 1;
                        }
                    }
                }
                counts[((size_t)r * n_motifs + m) * 2 + st] = cnt;
            }
        }
    }
    totals[0] = n_hits; totals[1] = n_near; totals[2] = ksum; totals[3] = kxor; totals[4] = n_flip;
    return overflow;
}

/* BASELINE config c5 (batched restatement of smeagol/variant.py:7-47 + scan.py:8-37): for every SNV (pos, alt code) on
 * ONE row of integer codes and every PWM, the maximum fp32 window score over all windows that cover pos and lie inside
 * the row, both strands, for the reference bases and with the base at pos replaced by alt.  Same fixed summation order
 * as orc_scan ("kernel" order).  ref_max / alt_max: [n_snv][n_motifs]; -inf where no window fits. */
void orc_variant_windows(const uint8_t* codes, int64_t L, const float* tab, const int32_t* widths, const int64_t* tab_off,
                         int n_motifs, const int64_t* snv_pos, const uint8_t* snv_alt, int64_t n_snv, float* ref_max,
                         float* alt_max)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t v = 0; v < n_snv; ++v) {
        const int64_t pos = snv_pos[v];
        for (int m = 0; m < n_motifs; ++m) {
            const int W = widths[m];
            const float* T = tab + tab_off[m] * 4;
            float rmax = -INFINITY, amax = -INFINITY;
            const int64_t s_lo = pos - W + 1 > 0 ? pos - W + 1 : 0;
            const int64_t s_hi = pos < L - W ? pos : L - W;
            for (int64_t s = s_lo; s <= s_hi; ++s)
                for (int st = 0; st < 2; ++st) {
                    float ra = 0.f, aa = 0.f;
                    for (int k = 0; k < W; ++k) {
                        const float* tr = T + (st ? (W - 1 - k) : k) * 4;
                        const float t0 = st ? tr[3] : tr[0], t1 = st ? tr[2] : tr[1], t2 = st ? tr[1] : tr[2],
                                    t3 = st ? tr[0] : tr[3];
                        const float* e = EMB[codes[s + k]];
                        const float* ea = (s + k == pos) ? EMB[snv_alt[v]] : e;
                        const float tr_ref = ((e[0] * t0 + e[1] * t1) + e[2] * t2) + e[3] * t3;
                        const float tr_alt = ((ea[0] * t0 + ea[1] * t1) + ea[2] * t2) + ea[3] * t3;
                        ra = k ? ra + tr_ref : tr_ref;
                        aa = k ? aa + tr_alt : tr_alt;
                    }
                    if (ra > rmax) rmax = ra;
                    if (aa > amax) amax = aa;
                }
            ref_max[v * n_motifs + m] = rmax;
            alt_max[v * n_motifs + m] = amax;
        }
    }
}
