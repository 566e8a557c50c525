// irn.cpp -- one InceptionResNet block (autoencoder.py:52-57) per C call.
//
//   out = cat( conv0_1(relu(conv0_0(x))),  conv1_2(relu(conv1_1(relu(conv1_0(x))))) ) + x
//
// The block is five layers (three k=3, two k=1) plus up to three format conversions; issuing them from Python costs
// ~17 us of interpreter time per launch (profiles/r01_host_profile.txt), which is what bounds the frame rate once
// two frames are in flight.  This entry point issues the same kernels in the same order from C: the caller resolves,
// once per layer, WHICH kernel serves it (route) and hands over the packed weights; nothing is computed differently.
#include <cstdint>

#include "../../include/pcgc.h"

namespace pcgc {
void set_error(const char *fmt, ...);
}

namespace {

inline bool h2_route(int r) { return r == PCGC_ROUTE_H2_GATHER || r == PCGC_ROUTE_H2_OCTET || r == PCGC_ROUTE_WIDE; }

// one k=3 layer of the block through the kernel its route names
int run_k3(const pcgc_irn_args *a, int i, const float *in_f, const uint32_t *in_h, int in_ld, int cin, int cout, const float *residual,
           int res_ld, float *out, int out_ld, uint32_t *out_h2, int out_h2_ld, int flags, void *stream) {
    const int64_t n = a->n, n_par = a->n / 8;
    switch (a->route[i]) {
        case PCGC_ROUTE_H2_GATHER:
            return pcgc_conv_k3_h2_fwd(in_h, in_ld, a->nbr, n, (const uint32_t *)a->w3[i], a->inv_scale[i], a->b3[i], cin, cout, residual,
                                       res_ld, out, out_ld, out_h2, out_h2_ld, flags, a->overflow, stream);
        case PCGC_ROUTE_WIDE:
            return pcgc_conv_k3_wide_fwd(in_h, in_ld, a->nbr, n, a->w3[i], a->inv_scale[i], a->b3[i], cin, cout, residual, res_ld, out,
                                         out_ld, out_h2, out_h2_ld, flags, a->overflow, stream);
        case PCGC_ROUTE_H2_OCTET:
            return pcgc_conv_k3_octet_h2_fwd(in_h, in_ld, a->parent_nbr, n_par, (const uint32_t *)a->w3[i], a->inv_scale[i], a->b3[i], cin,
                                             cout, residual, res_ld, out, out_ld, out_h2, out_h2_ld, flags, a->overflow, stream);
        case PCGC_ROUTE_TF32_GATHER:
            return pcgc_conv_k3_fwd_packed(in_f, in_ld, a->nbr, n, (const float *)a->w3[i], a->b3[i], cin, cout, residual, res_ld, out,
                                           out_ld, flags, stream);
        case PCGC_ROUTE_TF32_OCTET:
            return pcgc_conv_k3_octet_fwd(in_f, in_ld, a->parent_nbr, n_par, (const float *)a->w3[i], a->b3[i], cin, cout, residual,
                                          res_ld, out, out_ld, flags, stream);
        case PCGC_ROUTE_FP32:
            return pcgc_conv_k3_fwd(in_f, in_ld, a->nbr, n, (const float *)a->w3[i], a->b3[i], cin, cout, residual, res_ld, out, out_ld,
                                    flags, stream);
        default:
            pcgc::set_error("pcgc_irn_fwd: unknown route %d", a->route[i]);
            return PCGC_ERR_INVALID;
    }
}

}  // namespace

extern "C" {

size_t pcgc_irn_ws_bytes(int64_t n, int32_t c) { return (size_t)5 * (size_t)(n < 0 ? 0 : n) * (size_t)(c / 4) * 4 + 256; }

int pcgc_irn_fwd(const pcgc_irn_args *a, void *stream) {
    if (!a || a->n < 0 || a->c < 16 || a->c % 16 != 0) {
        pcgc::set_error("pcgc_irn_fwd: bad arguments");
        return PCGC_ERR_INVALID;
    }
    if (a->n == 0) return PCGC_OK;
    const int c = a->c, h = c / 2, q = c / 4;
    const bool want_h2 = a->out_h2 != nullptr;
    for (int i = 0; i < 3; ++i) {
        const bool needs_child_map = a->route[i] == PCGC_ROUTE_H2_GATHER || a->route[i] == PCGC_ROUTE_TF32_GATHER || a->route[i] == PCGC_ROUTE_FP32 ||
                                     a->route[i] == PCGC_ROUTE_WIDE;
        if ((needs_child_map && !a->nbr) || (!needs_child_map && (!a->parent_nbr || a->n % 8 != 0)) || !a->w3[i]) {
            pcgc::set_error("pcgc_irn_fwd: layer %d: kernel map or weights missing for route %d", i, a->route[i]);
            return PCGC_ERR_INVALID;
        }
    }
    if (!a->x || !a->out || !a->ws || a->ws_bytes < pcgc_irn_ws_bytes(a->n, c) || (!a->w1[0] && !(a->reserved & PCGC_IRN_MERGED_FIRST)) ||
        !a->w1[1]) {
        pcgc::set_error("pcgc_irn_fwd: null pointer or workspace too small");
        return PCGC_ERR_INVALID;
    }
    if (h2_route(a->route[0]) && !a->x_h2) {
        pcgc::set_error("pcgc_irn_fwd: conv0_0 runs on an h2 kernel but x_h2 is NULL");
        return PCGC_ERR_INVALID;
    }
    // workspace: a (fp32, h2), b (fp32, h2), cc (fp32); rows of q values, 256-byte aligned base
    char *base = (char *)(((uintptr_t)a->ws + 255) & ~(uintptr_t)255);
    const size_t plane = (size_t)a->n * q * 4;
    float *a_f = (float *)base;
    uint32_t *a_h = (uint32_t *)(base + plane);
    float *b_f = (float *)(base + 2 * plane);
    uint32_t *b_h = (uint32_t *)(base + 3 * plane);
    float *cc_f = (float *)(base + 4 * plane);
    int rc;

    // ---- merged first layers (PCGC_IRN_MERGED_FIRST): conv0_0 and conv1_0 read the same x and are both followed by a ReLU, so
    // they run as ONE k=3 convolution c -> 2q whose output channels q..2q-1 carry conv1_0's k=1 weights at the centre offset
    // (zero elsewhere).  With q = 4 the second half of the 8-wide MMA tile was idle anyway: conv1_0 costs nothing, its own
    // launch and its pass over x disappear.  The h2 result [n][2q] feeds conv0_1 (words 0..q-1) and conv1_1 (words q..2q-1).
    const bool merged = a->reserved & PCGC_IRN_MERGED_FIRST;
    if (merged && !(h2_route(a->route[0]) && h2_route(a->route[1]) && h2_route(a->route[2]) && q % 4 == 0)) {
        pcgc::set_error("pcgc_irn_fwd: merged first layers need h2 routes for all three k=3 layers and c %% 16 == 0");
        return PCGC_ERR_INVALID;
    }
    uint32_t *ab_h = a_h;                                        // merged: [n][2q] over the a_h and b_f planes
    const int ab_ld = merged ? 2 * q : q;
    if (merged) {
        rc = run_k3(a, 0, a->x, a->x_h2, a->x_h2_ld, c, 2 * q, nullptr, 0, nullptr, 0, ab_h, ab_ld, PCGC_EPI_RELU, stream);
        if (rc) return rc;
        b_h = ab_h + q;
    }

    if (a->reserved & PCGC_IRN_DUAL_SECOND) {                     // both second-stage branches in one kernel: two launches per block
        if (!merged || c != 16 || a->route[1] != PCGC_ROUTE_H2_OCTET || a->route[2] != PCGC_ROUTE_H2_OCTET) {
            pcgc::set_error("pcgc_irn_fwd: dual second stage needs merged first layers, c = 16 and the full-octet h2 routes");
            return PCGC_ERR_INVALID;
        }
        return pcgc_irn16_second_stage_fwd(ab_h, ab_ld, a->parent_nbr, a->n / 8, (const uint32_t *)a->w3[1], a->inv_scale[1], a->b3[1],
                                           (const uint32_t *)a->w3[2], a->inv_scale[2], a->b3[2], a->w1[1], a->b1[1], a->x, a->x_ld, a->out,
                                           a->out_ld, a->out_h2, a->out_h2_ld, a->overflow, stream);
    }

    // ---- branch 0: conv0_0 (c -> q, ReLU), conv0_1 (q -> h, + x[:, :h])
    const bool a_needs_h2 = h2_route(a->route[1]), a_needs_f32 = !a_needs_h2;
    if (merged) {
    } else if (h2_route(a->route[0]) && q % 4 == 0) {
        rc = run_k3(a, 0, a->x, a->x_h2, a->x_h2_ld, c, q, nullptr, 0, a_needs_f32 ? a_f : nullptr, q, a_needs_h2 ? a_h : nullptr, q,
                    PCGC_EPI_RELU, stream);
        if (rc) return rc;
    } else {
        const bool h2r = h2_route(a->route[0]);
        rc = run_k3(a, 0, a->x, a->x_h2, h2r ? a->x_h2_ld : a->x_ld, c, q, nullptr, 0, a_f, q, nullptr, 0, PCGC_EPI_RELU, stream);
        if (rc) return rc;
        if (a_needs_h2) {
            rc = pcgc_split_h2(a_f, q, a->n, q, a_h, q, a->overflow, stream);
            if (rc) return rc;
        }
    }
    const bool first_half_h2 = want_h2 && h2_route(a->route[1]);
    rc = run_k3(a, 1, a_f, a_h, ab_ld, q, h, a->x, a->x_ld, a->out, a->out_ld, first_half_h2 ? a->out_h2 : nullptr, a->out_h2_ld, 0, stream);
    if (rc) return rc;

    // ---- branch 1: conv1_0 (k=1, c -> q, ReLU), conv1_1 (k=3, q -> q, ReLU), conv1_2 (k=1, q -> h, + x[:, h:])
    const bool b_needs_h2 = h2_route(a->route[2]);
    if (merged) {
    } else if (b_needs_h2 && pcgc_conv_h2out_supported(1, c, q)) {
        rc = pcgc_conv_k1_fwd_h2out(a->x, a->x_ld, a->n, a->w1[0], a->b1[0], c, q, nullptr, 0, b_f, q, b_h, q, PCGC_EPI_RELU, a->overflow, stream);
        if (rc) return rc;
    } else {
        rc = pcgc_conv_k1_fwd(a->x, a->x_ld, a->n, a->w1[0], a->b1[0], c, q, nullptr, 0, b_f, q, PCGC_EPI_RELU, stream);
        if (rc) return rc;
        if (b_needs_h2) {
            rc = pcgc_split_h2(b_f, q, a->n, q, b_h, q, a->overflow, stream);
            if (rc) return rc;
        }
    }
    bool second_half_h2 = false;
    if (a->reserved & PCGC_IRN_FUSED_TAIL) {                      // conv1_1 + ReLU + conv1_2 + residual in one kernel: cc is never written
        if (a->route[2] != PCGC_ROUTE_H2_OCTET || !pcgc_conv_k3_octet_h2_k1_supported(q, q, h)) {
            pcgc::set_error("pcgc_irn_fwd: fused tail needs the full-octet h2 route and a supported shape (%d -> %d -> %d)", q, q, h);
            return PCGC_ERR_INVALID;
        }
        rc = pcgc_conv_k3_octet_h2_k1_fwd(b_h, ab_ld, a->parent_nbr, a->n / 8, (const uint32_t *)a->w3[2], a->inv_scale[2], a->b3[2], q, q,
                                          a->w1[1], a->b1[1], h, a->x + h, a->x_ld, a->out + h, a->out_ld, want_h2 ? a->out_h2 + h : nullptr,
                                          a->out_h2_ld, a->overflow, stream);
        if (rc) return rc;
        second_half_h2 = want_h2;
    } else {
    rc = run_k3(a, 2, b_f, b_h, ab_ld, q, q, nullptr, 0, cc_f, q, nullptr, 0, PCGC_EPI_RELU, stream);
    if (rc) return rc;
    if (want_h2 && pcgc_conv_h2out_supported(1, q, h)) {
        rc = pcgc_conv_k1_fwd_h2out(cc_f, q, a->n, a->w1[1], a->b1[1], q, h, a->x + h, a->x_ld, a->out + h, a->out_ld, a->out_h2 + h,
                                    a->out_h2_ld, 0, a->overflow, stream);
        second_half_h2 = true;
    } else {
        rc = pcgc_conv_k1_fwd(cc_f, q, a->n, a->w1[1], a->b1[1], q, h, a->x + h, a->x_ld, a->out + h, a->out_ld, 0, stream);
    }
    if (rc) return rc;
    }

    // ---- the halves of the h2 copy that no producer wrote
    if (want_h2) {
        if (!first_half_h2 && !second_half_h2) return pcgc_split_h2(a->out, a->out_ld, a->n, c, a->out_h2, a->out_h2_ld, a->overflow, stream);
        if (!first_half_h2) rc = pcgc_split_h2(a->out, a->out_ld, a->n, h, a->out_h2, a->out_h2_ld, a->overflow, stream);
        if (rc) return rc;
        if (!second_half_h2) rc = pcgc_split_h2(a->out + h, a->out_ld, a->n, h, a->out_h2 + h, a->out_h2_ld, a->overflow, stream);
    }
    return rc;
}

}  // extern "C"
