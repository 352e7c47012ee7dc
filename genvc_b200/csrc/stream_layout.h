// Layout of the decode weight stream (host + device).
//
// The fused decode kernel runs G persistent CTAs (one per SM).  Every GEMV of a decode step
//   QKV  : y[3D] = LN1(x) . W_attn   (K = D)        PROJ : x += o . W_proj        (K = D)
//   FC   : u[4D] = gelu(LN2(x) . W_fc) (K = D)      PROJ2: x += u . W_proj2       (K = 4D)
//   HEAD : logits[V] = z . mel_head^T  (K = D)
// is split by OUTPUT COLUMN: CTA c owns columns [c*N/G, (c+1)*N/G) of each matrix and needs the
// full K extent of those columns.  The stream stores, for each CTA, exactly the bytes it will
// consume, in consumption order, so the CTA's TMA producer walks one contiguous region of HBM:
//
//   stream = [ CTA 0 | CTA 1 | ... | CTA G-1 ]
//   CTA c  = [ layer 0 | layer 1 | ... | layer L-1 | head ]
//   layer  = [ QKV cols | PROJ cols | FC cols | PROJ2 cols ]
//   col    = K weights (the column of the [in,out] Conv1D matrix, i.e. K-contiguous) + bias + 3 pad
//
// so a column is (K+4) floats = a multiple of 16 bytes (bulk-copy granularity), and the bias add is
// folded into the dot product's initial value.  Total size = sum_p N_p (K_p + 4) per layer.
#pragma once
#include <stdint.h>

#ifndef GV_HD
#ifdef __CUDACC__
#define GV_HD __host__ __device__ __forceinline__
#else
#define GV_HD inline
#endif
#endif

namespace gv {

enum { PH_QKV = 0, PH_PROJ = 1, PH_FC = 2, PH_PROJ2 = 3, PH_HEAD = 4 };

struct StreamDims {
    int L, D, V, G;
};

GV_HD int ph_N(const StreamDims& s, int ph) {
    switch (ph) {
        case PH_QKV: return 3 * s.D;
        case PH_PROJ: return s.D;
        case PH_FC: return 4 * s.D;
        case PH_PROJ2: return s.D;
        default: return s.V;
    }
}
GV_HD int ph_K(const StreamDims& s, int ph) { return ph == PH_PROJ2 ? 4 * s.D : s.D; }
GV_HD long long col_begin(int N, int c, int G) { return ((long long)c * N) / G; }
GV_HD int col_owner(int N, int n, int G) { return (int)((((long long)n + 1) * G - 1) / N); }
GV_HD int ph_cols(const StreamDims& s, int ph, int c) {
    const int N = ph_N(s, ph);
    return (int)(col_begin(N, c + 1, s.G) - col_begin(N, c, s.G));
}
// floats of one layer of CTA c
GV_HD long long cta_layer_floats(const StreamDims& s, int c) {
    long long t = 0;
    for (int ph = PH_QKV; ph <= PH_PROJ2; ++ph) t += (long long)ph_cols(s, ph, c) * (ph_K(s, ph) + 4);
    return t;
}
// float offset of CTA c's region
GV_HD long long cta_base(const StreamDims& s, int c) {
    long long per_layer = 0;
    for (int ph = PH_QKV; ph <= PH_PROJ2; ++ph) per_layer += col_begin(ph_N(s, ph), c, s.G) * (ph_K(s, ph) + 4);
    return (long long)s.L * per_layer + col_begin(s.V, c, s.G) * (s.D + 4);
}
// float offset (inside CTA c's layer block) where phase `ph` starts
GV_HD long long ph_offset_in_layer(const StreamDims& s, int ph, int c) {
    long long t = 0;
    for (int q = PH_QKV; q < ph; ++q) t += (long long)ph_cols(s, q, c) * (ph_K(s, q) + 4);
    return t;
}
GV_HD long long stream_total_floats(const StreamDims& s) { return cta_base(s, s.G); }

// tile geometry of the shared-memory ring: a tile is 8 columns of a K = D matrix (one per consumer
// warp) or 2 columns of the K = 4D matrix — 32 KB at D = 1024; GV_MEGA_NSLOT tiles are in flight
#define GV_MEGA_NSLOT 6
GV_HD int slot_floats(int D) { return 8 * D + 32; }
GV_HD int tile_cols(int ph) { return ph == PH_PROJ2 ? 2 : 8; }

}  // namespace gv
