// Layout of the decode weight stream (host + device).
//
// The fused decode kernel runs G persistent CTAs (one per SM).  The four matrices of a block and
// the logits head are cut into UNITS of D weights (+ bias + 3 pad floats = D + 4 floats, a
// multiple of 16 bytes, the bulk-copy granularity):
//   QKV  : y[3D] = LN1(x) . W_attn      unit = one output column of W_attn  [D,3D]   (K = D)
//   PROJ : x1 = x + o . W_proj          unit = one output column of W_proj  [D,D]    (K = D)
//   FC   : u[4D] = gelu(LN2(x1) . W_fc) unit = one output column of W_fc    [D,4D]   (K = D)
//   P2   : x2 = x1 + u . W_proj2        unit = one ROW k of W_proj2 [4D,D]: the contribution of u_k
//                                              to all D outputs (the CTA that computed u_k owns row k,
//                                              so u never crosses CTAs; partial sums are reduced)
//   HEAD : logits[V] = z . mel_head^T   unit = one row of mel_head [V,D]            (K = D)
// CTA c owns units [c*N/G, (c+1)*N/G) of each matrix (FC and P2 share the range).  The stream
// stores, for each CTA, exactly the bytes it will consume, in consumption order, so the CTA's TMA
// producer walks one contiguous region of HBM:
//
//   stream = [ CTA 0 | CTA 1 | ... | CTA G-1 ]
//   CTA c  = [ layer 0 | layer 1 | ... | layer L-1 | head ]
//   layer  = [ QKV units | PROJ units | FC units | P2 units ]
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

enum { PH_QKV = 0, PH_PROJ = 1, PH_FC = 2, PH_P2 = 3, PH_HEAD = 4 };

struct StreamDims {
    int L, D, V, G;
};

// shared-memory ring of the fused kernel: GV_MEGA_NSLOT slots of GV_MEGA_UPT units each
#define GV_MEGA_NSLOT 12
#define GV_MEGA_UPT 4

GV_HD int unit_floats(int D) { return D + 4; }
GV_HD int slot_floats(int D) { return GV_MEGA_UPT * (D + 4); }

GV_HD int ph_N(const StreamDims& s, int ph) {
    switch (ph) {
        case PH_QKV: return 3 * s.D;
        case PH_PROJ: return s.D;
        case PH_FC: return 4 * s.D;
        case PH_P2: return 4 * s.D;
        default: return s.V;
    }
}
GV_HD long long col_begin(int N, int c, int G) { return ((long long)c * N) / G; }
GV_HD int col_owner(int N, int n, int G) { return (int)((((long long)n + 1) * G - 1) / N); }
GV_HD int ph_units(const StreamDims& s, int ph, int c) {
    const int N = ph_N(s, ph);
    return (int)(col_begin(N, c + 1, s.G) - col_begin(N, c, s.G));
}
// floats of one layer of CTA c
GV_HD long long cta_layer_floats(const StreamDims& s, int c) {
    long long t = 0;
    for (int ph = PH_QKV; ph <= PH_P2; ++ph) t += (long long)ph_units(s, ph, c) * unit_floats(s.D);
    return t;
}
// float offset of CTA c's region
GV_HD long long cta_base(const StreamDims& s, int c) {
    long long per_layer = 0;
    for (int ph = PH_QKV; ph <= PH_P2; ++ph) per_layer += col_begin(ph_N(s, ph), c, s.G) * unit_floats(s.D);
    return (long long)s.L * per_layer + col_begin(s.V, c, s.G) * unit_floats(s.D);
}
// float offset (inside CTA c's layer block) where phase `ph` starts
GV_HD long long ph_offset_in_layer(const StreamDims& s, int ph, int c) {
    long long t = 0;
    for (int q = PH_QKV; q < ph; ++q) t += (long long)ph_units(s, q, c) * unit_floats(s.D);
    return t;
}
GV_HD long long stream_total_floats(const StreamDims& s) { return cta_base(s, s.G); }

// attention work split of a decode step over S keys: items = H * nsplit, item = h * nsplit + split,
// split covers keys [split*chunk, min(S, (split+1)*chunk))
GV_HD int att_chunk(int S) {
    int c = (S + 7) / 8;
    if (c < 32) c = 32;
    return (c + 15) / 16 * 16;
}
GV_HD int att_nsplit(int S) { return (S + att_chunk(S) - 1) / att_chunk(S); }
// batched kernel: items = rows * H * nsplit must fit the grid (BH = rows * H <= G): at most G / BH key ranges per
// (row, head); chunk = keys per range (multiple of 16), actual nsplit = ceil(S / chunk)
GV_HD int att_nsplit_b(int S, int BH, int G) {
    const int ns = att_nsplit(S), cap = G / BH > 1 ? G / BH : 1;
    return ns < cap ? ns : cap;
}
GV_HD int att_chunk_b(int S, int nsplit) {
    int c = (S + nsplit - 1) / nsplit;
    if (c < 32) c = 32;
    return (c + 15) / 16 * 16;
}

}  // namespace gv
