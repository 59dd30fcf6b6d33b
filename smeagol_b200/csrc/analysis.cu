// SURVEY 8(f) rank 4 on the device: PWM similarity (smeagol/matrices.py:356-433 ncorr / pairwise_ncorrs, the only other
// super-linear numeric loop of the library: O(M^2 x alignments x 4W)) and window counting straight from the device hit
// list (smeagol/scan.py:381-423 _count_in_window / count_in_sliding_windows).
#include <string>

#include "common.cuh"

int smg_fail_from(int code, const std::string& msg);
void smg_count_launch();

namespace smg {

// np.corrcoef(x, y)[0, 1] of two aligned w x 4 blocks (matrices.py:263-287): covariance over the product of the standard
// deviations (ddof = 1 cancels but is kept so that the roundings match), clipped to [-1, 1]
__device__ double block_pearson(const double* X, const double* Y, const int n, const bool flipX, const bool flipY)
{
    // flip = the reference's reverse_complement AS WRITTEN (matrices.py:238-243): columns swapped, rows NOT reversed
    double mx = 0.0, my = 0.0;
    for (int i = 0; i < n; ++i) {
        mx += X[flipX ? ((i & ~3) | (3 - (i & 3))) : i];
        my += Y[flipY ? ((i & ~3) | (3 - (i & 3))) : i];
    }
    mx /= n;
    my /= n;
    double sxy = 0.0, sxx = 0.0, syy = 0.0;
    for (int i = 0; i < n; ++i) {
        const double dx = X[flipX ? ((i & ~3) | (3 - (i & 3))) : i] - mx;
        const double dy = Y[flipY ? ((i & ~3) | (3 - (i & 3))) : i] - my;
        sxy += dx * dy;
        sxx += dx * dx;
        syy += dy * dy;
    }
    const double n1 = (double)(n - 1);
    double c = (sxy / n1) / sqrt(sxx / n1) / sqrt(syy / n1);
    if (c > 1.0) c = 1.0;
    if (c < -1.0) c = -1.0;
    return c;
}

// ncorr of two matrices (rows x 4, row-major) with optional column flips; Python max() semantics over the alignments
__device__ double ncorr_dev(const double* A, int La, const bool flipA, const double* B, int Lb, const bool flipB)
{
    const double* X = A;
    const double* Y = B;
    int Lx = La, Ly = Lb;
    bool fx = flipA, fy = flipB;
    if (Lb < La) { X = B; Y = A; Lx = Lb; Ly = La; fx = flipB; fy = flipA; }  // order_pms: X is the narrower one
    const int mo = Lx < 3 ? Lx : 3;  // align_pms is always called with min_overlap=None (matrices.py:372)
    double best = 0.0;
    bool first = true;
    for (int i = mo - Lx; i <= Ly - mo; ++i) {
        const int xs = i < 0 ? -i : 0, xe = Lx < Ly - i ? Lx : Ly - i;
        const int ys = i > 0 ? i : 0, ye = Ly < i + Lx ? Ly : i + Lx;
        const int w = ye - ys;
        (void)xe;
        const double v = block_pearson(X + xs * 4, Y + ys * 4, 4 * w, fx, fy) * (double)w / (double)(Lx + Ly - w);
        if (first) { best = v; first = false; }
        else if (v > best) best = v;
    }
    return best;
}

__global__ void __launch_bounds__(128) pairwise_ncorrs_kernel(const double* __restrict__ mats, const int32_t* __restrict__ widths,
                                                              const int64_t* __restrict__ offs, const int n, double* __restrict__ out)
{
    const long long n_pairs = (long long)n * (n - 1) / 2;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_pairs; t += (long long)gridDim.x * blockDim.x) {
        // pair index -> (i, j), i < j, row-major over the strict upper triangle
        long long i = (long long)((2.0 * n - 1.0 - sqrt((2.0 * n - 1.0) * (2.0 * n - 1.0) - 8.0 * (double)t)) / 2.0);
        while (i * (2 * n - i - 1) / 2 > t) --i;
        while ((i + 1) * (2 * n - i - 2) / 2 <= t) ++i;
        const long long j = t - i * (2 * n - i - 1) / 2 + i + 1;
        const double* A = mats + offs[i] * 4;
        const double* B = mats + offs[j] * 4;
        const int La = widths[i], Lb = widths[j];
        // max over [A,B], [rc A, B], [A, rc B], [rc A, rc B] (matrices.py:417-425), Python max() semantics
        double best = ncorr_dev(A, La, false, B, Lb, false);
        double v = ncorr_dev(A, La, true, B, Lb, false);
        if (v > best) best = v;
        v = ncorr_dev(A, La, false, B, Lb, true);
        if (v > best) best = v;
        v = ncorr_dev(A, La, true, B, Lb, true);
        if (v > best) best = v;
        out[i * n + j] = best;
        out[j * n + i] = best;
    }
    for (long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x; d < n; d += (long long)gridDim.x * blockDim.x)
        out[d * n + d] = 1.0;
}

// every hit whose PWM is selected adds one to each tiling window [k * shift, k * shift + width) of its record that holds
// its start (scan.py:393-401: sites.id == window.id, window.start <= start < window.end)
__global__ void __launch_bounds__(256) window_counts_kernel(const unsigned long long* __restrict__ keys, const unsigned long long n_hits,
                                                            const int pos_bits, const int motif_bits, const uint8_t* __restrict__ motif_sel,
                                                            const int32_t* __restrict__ row_rec, const int64_t* __restrict__ win_off,
                                                            const long long width, const long long shift, int* __restrict__ counts)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_hits;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        const int motif = (int)(k & ((1ull << motif_bits) - 1ull));
        if (!motif_sel[motif]) continue;
        const long long pos = (long long)((k >> motif_bits) & ((1ull << pos_bits) - 1ull));
        const int rec = row_rec[k >> (pos_bits + motif_bits)];
        if (rec < 0) continue;
        const long long n_win = win_off[rec + 1] - win_off[rec];
        long long k_lo = pos - width + 1 > 0 ? (pos - width + shift) / shift : 0;  // ceil((pos - width + 1) / shift)
        long long k_hi = pos / shift;
        if (k_hi > n_win - 1) k_hi = n_win - 1;
        for (long long w = k_lo; w <= k_hi; ++w) atomicAdd(counts + win_off[rec] + w, 1);
    }
}

}  // namespace smg

extern "C" {

int smg_pairwise_ncorrs(const double* d_mats, const int32_t* d_widths, const int64_t* d_offs, int32_t n, double* d_out, void* stream)
{
    if (!d_mats || !d_widths || !d_offs || n <= 0 || !d_out) return smg_fail_from(SMG_ERR_INVALID, "smg_pairwise_ncorrs: bad argument");
    const long long n_pairs = (long long)n * (n - 1) / 2;
    const unsigned blocks = (unsigned)std::max<long long>(1, std::min<long long>((std::max<long long>(n_pairs, n) + 127) / 128, 148 * 32));
    smg::pairwise_ncorrs_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(d_mats, d_widths, d_offs, n, d_out);
    smg_count_launch();
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return smg_fail_from(SMG_ERR_CUDA, std::string("pairwise_ncorrs_kernel: ") + cudaGetErrorString(e));
    return SMG_OK;
}

int smg_window_counts(const uint64_t* d_keys, uint64_t n_hits, int pos_bits, int motif_bits, const uint8_t* d_motif_sel,
                      const int32_t* d_row_rec, const int64_t* d_win_off, int64_t width, int64_t shift, int32_t* d_counts, void* stream)
{
    if (pos_bits <= 0 || motif_bits <= 0 || pos_bits + motif_bits > 63 || !d_motif_sel || !d_row_rec || !d_win_off || width <= 0 ||
        shift <= 0 || !d_counts)
        return smg_fail_from(SMG_ERR_INVALID, "smg_window_counts: bad argument");
    if (n_hits == 0) return SMG_OK;
    if (!d_keys) return smg_fail_from(SMG_ERR_INVALID, "smg_window_counts: hit list missing");
    const unsigned blocks = (unsigned)std::min<uint64_t>((n_hits + 255) / 256, 148 * 16);
    smg::window_counts_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(d_keys), n_hits,
                                                                        pos_bits, motif_bits, d_motif_sel, d_row_rec, d_win_off, width,
                                                                        shift, d_counts);
    smg_count_launch();
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return smg_fail_from(SMG_ERR_CUDA, std::string("window_counts_kernel: ") + cudaGetErrorString(e));
    return SMG_OK;
}

}  // extern "C"
