// pcvae_build_weight_images: the hi / lo tf32 operand images of every weight matrix the four row-tile training kernels
// keep in shared memory, written to global memory in exactly the order each kernel lays them out, so that a CTA fetches
// its block with bulk async copies instead of building it (pcvae_tc_tile.cuh: fetch_images).  One launch, one thread per
// 16-byte chunk of an image; the values are those image_linear / image_linear_T produce (tests compare steps run with
// and without the prebuilt images bit for bit).
//
// Image of a layer: [chunk c of 4 reduction indices][image row r][4], hi at `dst`, lo = x - trunc_tf32(x) at `dst + lo`.
//   forward    (image_linear):   row r = output n, reduction index k = input: W[n][k]; column K holds the bias; with `one`
//                                row N has a 1 at column K (regenerates the constant-1 feature for the next layer)
//   transposed (image_linear_T): row r = input k, reduction index = output n: W[n][k]; no bias
// Image rows map to weight rows through up to two ranges (the encoder's last layer puts mean at rows [0, 10) and logvar at
// [16, 26)); every other entry is zero.
#include "pcvae_tc.cuh"
#include "pcvae_train.cuh"

namespace pcvae {

struct ImgDesc {
    int dst, lo;            // float offsets in the output buffer: hi image, distance to the lo image
    int C, nrows;           // chunks, image rows
    int W, b;               // offsets in theta (b < 0: none)
    int N, K;               // nn.Linear [N][K]
    int transposed, one;
    int r0[2], cnt[2], src[2];   // image rows [r0, r0 + cnt) <- weight rows (forward) starting at src
    int first;              // index of the image's first chunk in the launch
};
constexpr int MAX_IMG = 12;
struct ImgArgs {
    ImgDesc d[MAX_IMG];
    int n, total;
    const float* theta;
    float* out;
};

__global__ void __launch_bounds__(256) k_build_weight_images(const ImgArgs a) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += gridDim.x * blockDim.x) {
        int j = 0;
        while (j + 1 < a.n && i >= a.d[j + 1].first) ++j;
        const ImgDesc& d = a.d[j];
        const int e = i - d.first, c = e / d.nrows, r = e - c * d.nrows;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (d.transposed) {                     // row r = input index, reduction = output index 4c + q
            if (r < d.K) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (4 * c + q < d.N) v[q] = __ldg(a.theta + d.W + (long)(4 * c + q) * d.K + r);
            }
        } else {
            int n = -1;                         // weight row of image row r
            if (r >= d.r0[0] && r < d.r0[0] + d.cnt[0]) n = d.src[0] + (r - d.r0[0]);
            else if (r >= d.r0[1] && r < d.r0[1] + d.cnt[1]) n = d.src[1] + (r - d.r0[1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = 4 * c + q;
                if (n >= 0) {
                    if (k < d.K) v[q] = __ldg(a.theta + d.W + (long)n * d.K + k);
                    else if (k == d.K && d.b >= 0) v[q] = __ldg(a.theta + d.b + n);
                } else if (d.one && r == d.N && k == d.K) {
                    v[q] = 1.0f;
                }
            }
        }
        float* hi = a.out + d.dst + (long)e * 4;
        *reinterpret_cast<float4*>(hi) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(hi + d.lo) = make_float4(tc::tf32_lo(v[0]), tc::tf32_lo(v[1]), tc::tf32_lo(v[2]), tc::tf32_lo(v[3]));
    }
}

namespace {
struct Planner {
    ImgArgs a{};
    int cursor = 0;         // float offset of the next image in the buffer
    void begin(long base) { cursor = (int)base; }
    // one layer = hi image followed by lo image, as the kernels carve them
    ImgDesc& add(int C, int nrows, int W, int b, int N, int K, bool transposed, bool one) {
        ImgDesc& d = a.d[a.n++];
        d.dst = cursor; d.lo = C * nrows * 4; d.C = C; d.nrows = nrows; d.W = W; d.b = b; d.N = N; d.K = K;
        d.transposed = transposed; d.one = one;
        d.r0[0] = 0; d.cnt[0] = transposed ? 0 : N; d.src[0] = 0;
        d.r0[1] = 0; d.cnt[1] = 0; d.src[1] = 0;
        d.first = a.total;
        a.total += C * nrows;
        cursor += 2 * C * nrows * 4;
        return d;
    }
};
}  // namespace

int build_weight_images_launch(const Layout& L, const Layout& Le, bool enc, bool dec, const WeightImages& w, const float* theta,
                               float* images, cudaStream_t st) {
    using namespace tc;
    Planner p;
    if (enc) {
        const int D = Le.D, C1 = ((D + 8) & ~7) / 4;
        p.begin(w.enc_fwd);                                                     // k_enc_fwd_tc: W1, W2, W3 (mean | logvar split)
        p.add(C1, E1_N, Le.W1, Le.b1, H1, D, false, true);
        p.add(E2_C, E2_N, Le.W2, Le.b2, H2, H1, false, true);
        ImgDesc& d3 = p.add(E3_C, E3_N, Le.W3, Le.b3, LAT2, H2, false, false);
        d3.r0[0] = 0;  d3.cnt[0] = LAT; d3.src[0] = 0;                          // mean rows -> image rows [0, 10)
        d3.r0[1] = 16; d3.cnt[1] = LAT; d3.src[1] = LAT;                        // logvar rows -> image rows [16, 26)
        p.begin(w.enc_bwd);                                                     // k_enc_bwd_tc: W3^T, W2^T
        p.add(Y3_C, Y3_N, Le.W3, -1, LAT2, H2, true, false);
        p.add(Y2_C, Y2_N, Le.W2, -1, H2, H1, true, false);
    }
    if (dec) {
        const int D = L.D, N6 = (D + 15) & ~15;
        p.begin(w.dec_fwd);                                                     // k_dec_fwd_tc: W4, W5, W6
        p.add(F4_C, F4_N, L.W4, L.b4, G1, LAT, false, true);
        p.add(F5_C, F5_N, L.W5, L.b5, G2, G1, false, true);
        p.add(F6_C, N6, L.W6, L.b6, D, G2, false, false);
        p.begin(w.dec_bwd);                                                     // k_dec_bwd_tc: W6^T, W5^T, W4^T
        p.add(X6_C, X6_N, L.W6, -1, D, G2, true, false);
        p.add(X5_C, X5_N, L.W5, -1, G2, G1, true, false);
        p.add(X4_C, X4_N, L.W4, -1, G1, LAT, true, false);
    }
    p.a.theta = theta;
    p.a.out = images;
    if (p.a.total == 0) return PCVAE_OK;
    k_build_weight_images<<<(p.a.total + 255) / 256, 256, 0, st>>>(p.a);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "build_weight_images: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

}  // namespace pcvae
