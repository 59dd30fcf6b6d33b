// Batched SNV scoring kernel (BASELINE config 5; batched restatement of smeagol/variant.py:7-47, scan.py:8-37).
//
// For every SNV (row, pos, alt) and every PWM: over all windows that cover `pos` on both strands and lie inside the
// sequence, the maximum reference score and the maximum alternate score.  Same machinery as the fused scan kernel:
// one warp per CTA, lane = PWM with both strands' tables in registers, the 2*WB-1 bases around the SNV are the
// warp-uniform operand, window sums advance as a FADD2 shift register (ascending-k order).  Two passes per SNV
// (reference bases, then the same bases with the SNV position replaced by the alternate base).
#pragma once
#include "scan_regtab.cuh"

namespace smg {

struct VariantParams {
    ScanParams sp;
    const int32_t* snv_row;
    const int32_t* snv_pos;
    const uint8_t* snv_alt;
    long long n_snv;
    float* ref_max;
    float* alt_max;
    unsigned int* slow_flags;  // [n_snv] set when the SNV needs the generic path (ambiguity codes in the region)
};

#define SMG_VSTEP(B, O)                                                                       \
    {                                                                                         \
        O = vadd(r[WB - 2], t.get(WB - 1, B));                                                \
        _Pragma("unroll") for (int k = WB - 2; k >= 1; --k) r[k] = vadd(r[k - 1], t.get(k, B)); \
        r[0] = t.get(0, B);                                                                   \
    }

template <int WB>
__global__ void __launch_bounds__(32) variant_regtab_kernel(const VariantParams vp)
{
    using V = float2;
    const ScanParams& p = vp.sp;
    const int lane = threadIdx.x;
    const int group = p.first_group + blockIdx.y;
    const int sub = blockIdx.z;  // PWM slot of the lane group handled by this CTA
    Tab<WB> t;
    t.load(p.group_tab + p.group_tab_off[group] + (size_t)sub * (WB * 4 * 32 * 2), lane);
    const int motif = p.lane_motif[(group * SMG_MAX_PWMS_PER_LANE + sub) * 32 + lane];
    const int wm = motif >= 0 ? p.motif_width[motif] : 1;
    const int rc_shift = WB - (motif >= 0 ? wm : WB);  // head-padded RC table: its window completes rc_shift steps earlier
    const float ninf = __int_as_float(0xff800000);
    constexpr int NSTEP = 2 * WB - 1;

    for (long long v = blockIdx.x; v < vp.n_snv; v += gridDim.x) {
        const smg_row_t row = p.rows[vp.snv_row[v]];
        const long long pos = vp.snv_pos[v];
        const int alt = vp.snv_alt[v];
        const long long q0 = pos - (WB - 1);  // first base fed (may be < 0: those windows are invalid anyway)
        // ---- gather the 2*WB-1 <= 63 bases into two 64-bit words (uniform); detect ambiguity codes in the region
        unsigned long long lo = 0, hi = 0;
        bool slow = (alt < 1 || alt > 5);
        if (pos < 0 || pos >= row.n_avail) {  // out of range: the generic kernel reports -inf
            if (lane == 0 && blockIdx.y == 0 && blockIdx.z == 0) vp.slow_flags[v] = 1u;
            continue;
        }
        {
            const long long qa = q0 < 0 ? 0 : q0;
            const long long wfirst = row.word_off + (qa >> 4);
            unsigned long long A = 0, B = 0, C = 0;
            unsigned amb = 0;
            uint32_t wds[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                wds[i] = p.packed[wfirst + i];
                amb |= (unsigned)p.amask[wfirst + i] << 0;  // any ambiguity in the touched words (conservative)
            }
            // bases beyond n_avail are padding (mask set) -- only matters if a valid window would touch them, which the
            // bounds below exclude; restrict the ambiguity test to bases inside the row
            const long long qend = q0 + NSTEP;  // exclusive
            if (amb) {
                for (long long q = qa; q < qend && q < row.n_avail; ++q)
                    if ((p.amask[row.word_off + (q >> 4)] >> (q & 15)) & 1) slow = true;
            }
            A = (unsigned long long)wds[0] | ((unsigned long long)wds[1] << 32);
            B = (unsigned long long)wds[2] | ((unsigned long long)wds[3] << 32);
            C = (unsigned long long)wds[4];
            const int sh = 2 * (int)(qa & 15);
            lo = sh ? ((A >> sh) | (B << (64 - sh))) : A;
            hi = sh ? ((B >> sh) | (C << (64 - sh))) : B;
            const int lead = (int)(qa - q0);  // bases before the start of the sequence: shift them in as 'A'
            if (lead > 0) {
                hi = lead >= 32 ? (lo << (2 * (lead - 32))) : ((hi << (2 * lead)) | (lo >> (64 - 2 * lead)));
                lo = lead >= 32 ? 0ull : (lo << (2 * lead));
            }
        }
        if (slow) {  // warp-uniform: leave this SNV to the generic kernel
            if (lane == 0 && blockIdx.y == 0 && blockIdx.z == 0) vp.slow_flags[v] = 1u;
            continue;
        }
        // per-lane range of steps whose completed window is valid: covers pos, starts >= 0, ends inside the sequence
        const long long smax64 = row.g_len - row.g_off - (long long)wm;
        long long imin = 2 * (WB - 1) - wm + 1;
        const long long need0 = 2 * (WB - 1) - pos;  // s_i >= 0
        if (need0 > imin) imin = need0;
        long long imax = smax64 - pos + 2 * (WB - 1);  // s_i <= smax
        if (imax > 2 * (WB - 1)) imax = 2 * (WB - 1);
        float res[2];
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            unsigned long long L0 = lo, H0 = hi;
            if (pass == 1) {  // substitute the alternate base at fed index WB-1
                const unsigned long long a2 = (unsigned long long)(alt == 5 ? 3 : alt - 1);
                constexpr int idx = WB - 1;
                if constexpr (idx < 32) L0 = (L0 & ~(3ull << (2 * idx))) | (a2 << (2 * idx));
                else H0 = (H0 & ~(3ull << (2 * (idx & 31)))) | (a2 << (2 * (idx & 31)));
            }
            V r[WB - 1];
#pragma unroll
            for (int k = 0; k < WB - 1; ++k) r[k] = V{};
            float mx = ninf;
#pragma unroll 1
            for (int i = 0; i < NSTEP; ++i) {
                V o;
                if (L0 & 2ull) {
                    if (L0 & 1ull) SMG_VSTEP(3, o) else SMG_VSTEP(2, o)
                } else {
                    if (L0 & 1ull) SMG_VSTEP(1, o) else SMG_VSTEP(0, o)
                }
                L0 = (L0 >> 2) | (H0 << 62);
                H0 >>= 2;
                if (i >= imin && i <= imax) mx = fmaxf(mx, o.x);
                if (i >= imin - rc_shift && i <= imax - rc_shift) mx = fmaxf(mx, o.y);
            }
            res[pass] = mx;
        }
        if (motif >= 0) {
            vp.ref_max[v * p.n_motifs + motif] = res[0];
            vp.alt_max[v * p.n_motifs + motif] = res[1];
        }
    }
}

template <int WB>
void launch_variant_regtab(const VariantParams& vp, int n_blocks, int n_groups, int nm, cudaStream_t stream)
{
    dim3 grid((unsigned)n_blocks, (unsigned)n_groups, (unsigned)nm);
    variant_regtab_kernel<WB><<<grid, 32, 0, stream>>>(vp);
}

}  // namespace smg
