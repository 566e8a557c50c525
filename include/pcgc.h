/* pcgc.h -- C ABI of libpcgc: the B200-native (sm_100a) hot path of the PCGCv2
 * multiscale sparse-conv point-cloud geometry codec.
 *
 * The reference (NJUVISION/PCGCv2) has no C ABI of its own: its hot path lives
 * in the third-party MinkowskiEngine / torchac packages, entered from Python.
 * Each entry point below names the reference call site whose native work it
 * replaces (file:line in the reference repository) and the SURVEY.md section 8
 * row it implements.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - no allocation inside: callers pass outputs and (where needed) a workspace
 *     whose size comes from the matching *_ws_bytes() query;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     calls are asynchronous on that stream unless stated otherwise;
 *   - return value: 0 = ok, <0 = error, text via pcgc_last_error() (thread
 *     local); no exception crosses the boundary;
 *   - coordinate KEYS: uint64 = (batch << 57) | morton3(x, y, z) of the
 *     coordinate divided by the tensor stride (x in the lowest bit of each
 *     triple), 19 bits per axis, batch < 127.  One level up the octree shifts
 *     the 57-bit Morton field right by 3 (batch bits stay); the child index k
 *     (= ix + 2*iy + 4*iz, the reference's kernel index for k=2 kernels) is
 *     key & 7.  PCGC_EMPTY_KEY marks a free hash slot.
 */
#ifndef PCGC_H_
#define PCGC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCGC_OK 0
#define PCGC_ERR_INVALID (-1)
#define PCGC_ERR_CUDA (-2)
#define PCGC_ERR_WORKSPACE (-3)
#define PCGC_ERR_RANGE (-4)
#define PCGC_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define PCGC_MAX_COORD ((1 << 19) - 1)
#define PCGC_MAX_BATCH 126

/* epilogue flags of the convolution entry points */
#define PCGC_EPI_RELU 1       /* out = max(out, 0) after bias (+ residual) */
#define PCGC_TILES_CHUNKED 256 /* full-octet kernels: every CTA walks one contiguous run of tiles (L1 reuse of shared halo faces)
                                * instead of the strided order; set by the library itself (PCGC_OCTET_TILE_ORDER=0 disables) */

int pcgc_version(void);
const char *pcgc_last_error(void);
/* number of kernels libpcgc has launched in this process (bench.py's gpu_launches) */
uint64_t pcgc_launch_count(void);

/* ---- coordinates ------------------- */

/* a1  ME.SparseTensor(features, coordinates, tensor_stride) -- data_utils.py:96,108,116,
 * coder.py:102.  coords int32 [n,4] = (b,x,y,z), every spatial entry a multiple of
 * tensor_stride.  err_flag (device int32, caller zeroes it) is set to 1 on a
 * coordinate outside [0, PCGC_MAX_COORD*stride], a batch > PCGC_MAX_BATCH or a
 * non-multiple of the stride. */
int pcgc_pack_keys(const int32_t *coords, int64_t n, int32_t tensor_stride, uint64_t *keys,
                   int32_t *err_flag, void *stream);

/* the same from int32 [n,3] = (x,y,z) rows with one batch index for all rows (coder.py is
 * batch-1 by construction, coder.py:97,106): no padded [n,4] copy of the cloud is needed.
 * hint_bits > 0: the caller expects every coordinate / stride below 2^hint_bits and batch 0 (the --res of coder.py:196) and
 * will sort only the low 3 * hint_bits key bits; bit 1 of *err_flag is raised when a row breaks the hint (the caller then
 * repeats with the full key width), bit 0 as above. */
int pcgc_pack_keys3(const int32_t *coords3, int64_t n, int32_t tensor_stride, int32_t batch, int32_t hint_bits, uint64_t *keys,
                    int32_t *err_flag, void *stream);

/* a5  gather map of ME.MinkowskiConvolution(k=2, s=2) -- autoencoder.py:78,97,116: child_map
 * int32 [8][n_parents], entry [k][p] = row of child k (= key & 7 = ix + 2 iy + 4 iz) of parent p
 * or -1; parent_of int32 [n] comes from pcgc_stride_down. */
int pcgc_child_map_k2(const uint64_t *child_keys, const int32_t *parent_of, int64_t n, int64_t n_parents,
                      int32_t *child_map, void *stream);

/* scale_sparse_tensor(x, factor) -- data_utils.py:112-118 (coder.py:149-152,166-167): every
 * coordinate value v -> (int) round_half_even((float) v * factor); `count` = number of int32
 * values (rows x columns; pass the spatial columns only).  The duplicate rows the down-scaling
 * creates collapse at the following coordinate-map insertion (pcgc_hash_build / Codec.encode). */
int pcgc_scale_coords(const int32_t *coords, int64_t count, float factor, int32_t *out, void *stream);
int pcgc_unpack_keys(const uint64_t *keys, int64_t n, int32_t tensor_stride, int32_t *coords,
                     void *stream);

/* capacity (slots, a power of two >= 2n) of the hash table for n keys */
int64_t pcgc_hash_capacity(int64_t n);
/* a1  coordinate-map insertion.  Fills table_keys[cap] / table_vals[cap] (cleared here)
 * with key -> smallest row holding it; n_dup (device int32) receives the number of
 * rows whose key was already present (0 = all rows unique). */
int pcgc_hash_build(const uint64_t *keys, int64_t n, uint64_t *table_keys, int32_t *table_vals,
                    int64_t cap, int32_t *n_dup, void *stream);
/* first-seen flags for de-duplication (A.2): keep[i] = 1 iff row i is the
 * representative of its key in the table built by pcgc_hash_build. */
int pcgc_hash_keep_flags(const uint64_t *keys, int64_t n, const uint64_t *table_keys,
                         const int32_t *table_vals, int64_t cap, uint8_t *keep, void *stream);
/* a11 isin(data.C, gt.C) -- data_utils.py:63-75.  found[i] = 1 iff query key i is in the table. */
int pcgc_hash_contains(const uint64_t *query, int64_t n, const uint64_t *table_keys, int64_t cap,
                       uint8_t *found, void *stream);

/* a4  kernel map of a k=3 stride-1 convolution (coordinate_manager.kernel_map behind every
 * k=3 ME.MinkowskiConvolution in autoencoder.py:13,20,35,71,90,109,128,162,...).
 * nbr int32 [27][n] (offset-major): nbr[k*n + u] = row of coords[u] + offset_k or -1,
 * k = ix + 3*iy + 9*iz, offsets {-1,0,1} * tensor_stride, x fastest.  n_pairs (device
 * int64, may be NULL) receives the number of (in,out) pairs. */
int pcgc_kernel_map_k3(const uint64_t *keys, int64_t n, const uint64_t *table_keys,
                       const int32_t *table_vals, int64_t cap, int32_t *nbr,
                       unsigned long long *n_pairs, void *stream);

/* a5  output coordinate map + kernel map of ME.MinkowskiConvolution(kernel_size=2, stride=2)
 * -- autoencoder.py:78,97,116.  Parents = unique(parent(key)) in ascending key order.
 * Outputs: parent_keys[<=n], n_parents (device int32), child_rows int32 [n] (input rows
 * grouped by parent, ascending child index), child_off int32 [n_parents+1] (CSR offsets
 * into child_rows; caller provides n+1 entries), parent_of int32 [n] (may be NULL): parent
 * row of every input row.  keys_are_sorted != 0 promises ascending
 * keys (the order every map derived by this library has) and skips the radix sort. */
size_t pcgc_stride_down_ws_bytes(int64_t n);
int pcgc_stride_down(const uint64_t *keys, int64_t n, int32_t keys_are_sorted, uint64_t *parent_keys,
                     int32_t *n_parents, int32_t *child_rows, int32_t *child_off, int32_t *parent_of,
                     void *ws, size_t ws_bytes, void *stream);

/* a4  kernel map of a SORTED child set derived from the kernel map of its parent set -- no
 * hashing: neighbour (c + d) of child position c lies in parent neighbour floor((c+d)/2) at
 * child position (c+d)&1.  parent_info[p] = (first child row << 8) | occupancy byte, from
 * pcgc_parent_info; parent_of[i] from pcgc_stride_down.  parent_info == NULL selects the
 * full-octet mode (child set = generative up-sampling output, row 8*p + c; no tables read).
 * parent_nbr is the parent's map [27][n_parents]; nbr the child's [27][n]. */
int pcgc_parent_info(const uint64_t *child_keys, const int32_t *child_off, int64_t n_parents,
                     uint64_t *parent_info, void *stream);
int pcgc_kernel_map_k3_from_parent(const uint64_t *child_keys, const int32_t *parent_of,
                                   const uint64_t *parent_info, const int32_t *parent_nbr,
                                   int64_t n_parents, int64_t n, int32_t *nbr, void *stream);

/* a6  output coordinate map of ME.MinkowskiGenerativeConvolutionTranspose(k=2, s=2)
 * -- autoencoder.py:155,182,209: child_keys[8*i + k] = child(keys[i], k) (Morton field << 3 | k). */
int pcgc_upsample_keys(const uint64_t *keys, int64_t n, uint64_t *child_keys, void *stream);

/* generic helpers: exclusive scan of a uint8 mask and stable argsort of uint64 keys
 * (a10 sort_spare_tensor -- data_utils.py:91-101, coder.py:97-99). */
size_t pcgc_argsort_ws_bytes(int64_t n);
int pcgc_argsort_u64(const uint64_t *keys, int64_t n, int end_bit, uint64_t *keys_sorted,
                     int32_t *order, void *ws, size_t ws_bytes, void *stream);

/* ---- convolutions (rows a3, a5, a6, a7) -------------------------------------------------
 * Features are float32 row-major with an explicit leading dimension (ld, in floats) so a
 * layer can read/write a column slice of a wider tensor (ME.cat fusion): element (r, c) of
 * a tensor is base[r*ld + c].  weight is the reference parameter `kernel` [K][cin][cout]
 * (2-D [cin][cout] for K = 1), bias is `bias` [1][cout] or NULL.  residual (may be NULL)
 * is added before the optional ReLU: out = epi(conv + bias + residual). */

/* a3  ME.MinkowskiConvolution(kernel_size=3, stride=1).forward.  nbr from pcgc_kernel_map_k3. */
int pcgc_conv_k3_fwd(const float *in, int32_t in_ld, const int32_t *nbr, int64_t n, const float *weight,
                     const float *bias, int32_t cin, int32_t cout, const float *residual,
                     int32_t res_ld, float *out, int32_t out_ld, int32_t flags, void *stream);
/* a3 on the tensor cores (3xTF32, FP32-accurate): the same convolution with the weights
 * pre-packed ONCE per layer into mma B-fragment order.  pcgc_conv_k3_packed_floats returns the
 * packed size (0 = this cin x cout has no tensor-core kernel: cin in {8,16,32,64}, cout in
 * {1,4,8,16,32,64}); input rows must be 16-byte aligned. */
size_t pcgc_conv_k3_packed_floats(int32_t cin, int32_t cout);
int pcgc_conv_k3_pack_weights(const float *weight, int32_t cin, int32_t cout, float *packed, void *stream);
int pcgc_conv_k3_fwd_packed(const float *in, int32_t in_ld, const int32_t *nbr, int64_t n, const float *packed,
                            const float *bias, int32_t cin, int32_t cout, const float *residual,
                            int32_t res_ld, float *out, int32_t out_ld, int32_t flags, void *stream);
/* a3 on FULL-OCTET sets (every set the synthesis network convolves on is the 8-child expansion of a parent
 * set, ME.MinkowskiGenerativeConvolutionTranspose autoencoder.py:155,182,209: n = 8 * n_parents rows, row
 * 8*i + c = child c = cx + 2cy + 4cz of parent row i).  The 4x4x4 voxel halo of each octet is staged in shared
 * memory (cp.async, contiguous sibling runs) and addressed by the PARENT set's kernel map parent_nbr
 * [27][n_parents] alone -- the child-level map is not needed.  Same result as pcgc_conv_k3_fwd(_packed) on the
 * child map.  pcgc_conv_k3_octet_packed_floats returns 0 for shapes without such a kernel (cin in {4,8,16},
 * cout in {1,4,8,16}; cin 4: cout in {4,8}). */
size_t pcgc_conv_k3_octet_packed_floats(int32_t cin, int32_t cout);
int pcgc_conv_k3_octet_pack_weights(const float *weight, int32_t cin, int32_t cout, float *packed, void *stream);
int pcgc_conv_k3_octet_fwd(const float *in, int32_t in_ld, const int32_t *parent_nbr, int64_t n_parents,
                           const float *packed, const float *bias, int32_t cin, int32_t cout,
                           const float *residual, int32_t res_ld, float *out, int32_t out_ld, int32_t flags,
                           void *stream);
/* a3 over PRE-SPLIT half-precision features ("h2" format; conv_h2.cuh).  An h2 tensor holds x = hi + lo with
 * hi = f16(x), lo = f16(x - hi) (22 significand bits, 4 bytes per value like the fp32 it replaces; |x| < 65504):
 * uint32 [n][c] with leading dimension in 4-byte units, every group of four channels stored as the 16 bytes
 * {hi0 hi1 | hi2 hi3 | lo0 lo1 | lo2 lo3}, so that a gathered row is loaded straight into mma.sync.m16n8k16 f16
 * fragments and the hot loop has no split arithmetic (the 3xTF32 kernels above split every row once per
 * neighbour that gathers it).  Weights are packed once per layer after multiplication by `scale` (a power of two
 * that keeps the lo parts normal); pass inv_scale = 1/scale to the convolution.  The convolution writes fp32
 * (`out`, may be NULL) and/or h2 (`out_h2`, may be NULL; cout % 4 == 0) for the next k=3 layer; *overflow (may be
 * NULL) is set to 1 when a value written in h2 leaves the f16 range, so the caller can re-run on the fp32 path.
 * Same result as pcgc_conv_k3_fwd to ~1e-6 of max|out|.  packed_words returns 0 for shapes without a kernel
 * (cin in {16,32,64}). */
int pcgc_split_h2(const float *in, int32_t in_ld, int64_t n, int32_t c, uint32_t *out_h2, int32_t out_ld,
                  int32_t *overflow, void *stream);
int pcgc_join_h2(const uint32_t *in_h2, int32_t in_ld, int64_t n, int32_t c, float *out, int32_t out_ld, void *stream);
size_t pcgc_conv_k3_h2_packed_words(int32_t cin, int32_t cout);
int pcgc_conv_k3_h2_pack_weights(const float *weight, int32_t cin, int32_t cout, float scale, uint32_t *packed,
                                 void *stream);
int pcgc_conv_k3_h2_fwd(const uint32_t *in_h2, int32_t in_ld, const int32_t *nbr, int64_t n, const uint32_t *packed,
                        float inv_scale, const float *bias, int32_t cin, int32_t cout, const float *residual,
                        int32_t res_ld, float *out, int32_t out_ld, uint32_t *out_h2, int32_t out_h2_ld,
                        int32_t flags, int32_t *overflow, void *stream);
/* a3 on FULL-OCTET sets over h2 features: the halo staging of pcgc_conv_k3_octet_fwd feeding the f16 arithmetic of
 * pcgc_conv_k3_h2_fwd (same packed weights and scale as the latter).  cin = 16, cout in {1,4,8,16,32}; and cin = 4,
 * cout in {4,8} (full-octet only: one MMA carries hi*Whi + lo*Whi + hi*Wlo of a 4-channel row). */
int pcgc_conv_k3_octet_h2_supported(int32_t cin, int32_t cout);
int pcgc_conv_k3_octet_h2_fwd(const uint32_t *in_h2, int32_t in_ld, const int32_t *parent_nbr, int64_t n_parents,
                              const uint32_t *packed, float inv_scale, const float *bias, int32_t cin, int32_t cout,
                              const float *residual, int32_t res_ld, float *out, int32_t out_ld, uint32_t *out_h2,
                              int32_t out_h2_ld, int32_t flags, int32_t *overflow, void *stream);
/* a5 / a6 on the tensor cores over h2 features.
 * k=2 stride 2 (a5): the h2 gather kernel over the 8 child slots of every parent.  child_map int32 [8][n_parents]:
 * child_map[k][p] = input row of child k = in_key & 7 of parent p, or -1.  Shapes 16x32, 32x64, 64x32.
 * Transposed k=2 stride 2 (a6): ONE dense product [n_in, cin] x [cin, 8*cout] -- the [8*n_in, cout] child tensor is
 * the same memory as [n_in, 8*cout], so out / out_h2 must be contiguous (ld == cout).  bias8 = the bias repeated 8
 * times (8*cout floats).  pack_weights needs a float workspace of 8*cin*cout for the re-ordered kernel.
 * cin in {16,32,64}, cout <= 64. */
size_t pcgc_conv_k2s2_h2_packed_words(int32_t cin, int32_t cout);
int pcgc_conv_k2s2_h2_pack_weights(const float *weight, int32_t cin, int32_t cout, float scale, uint32_t *packed,
                                   void *stream);
int pcgc_conv_k2s2_h2_fwd(const uint32_t *in_h2, int32_t in_ld, const int32_t *child_map, int64_t n_parents,
                          const uint32_t *packed, float inv_scale, const float *bias, int32_t cin, int32_t cout,
                          float *out, int32_t out_ld, uint32_t *out_h2, int32_t out_h2_ld, int32_t flags,
                          int32_t *overflow, void *stream);
size_t pcgc_convT_k2s2_h2_packed_words(int32_t cin, int32_t cout);
int pcgc_convT_k2s2_h2_pack_weights(const float *weight, int32_t cin, int32_t cout, float scale, float *dense_ws,
                                    uint32_t *packed, void *stream);
int pcgc_convT_k2s2_h2_fwd(const uint32_t *in_h2, int32_t in_ld, int64_t n_in, const uint32_t *packed, float inv_scale,
                           const float *bias8, int32_t cin, int32_t cout, float *out, int32_t out_ld,
                           uint32_t *out_h2, int32_t out_h2_ld, int32_t flags, int32_t *overflow, void *stream);
/* a7  ME.MinkowskiConvolution(kernel_size=1) == F.mm(kernel) + bias. */
int pcgc_conv_k1_fwd(const float *in, int32_t in_ld, int64_t n, const float *weight, const float *bias,
                     int32_t cin, int32_t cout, const float *residual, int32_t res_ld, float *out,
                     int32_t out_ld, int32_t flags, void *stream);
/* a5  ME.MinkowskiConvolution(kernel_size=2, stride=2).forward: out[p] = b + sum over the
 * children j of p of in[child_rows[j]] @ W[child_k[j]], child_k = in_keys[row] & 7. */
int pcgc_conv_k2s2_fwd(const float *in, int32_t in_ld, const uint64_t *in_keys, const int32_t *child_rows,
                       const int32_t *child_off, int64_t n_parents, const float *weight, const float *bias,
                       int32_t cin, int32_t cout, float *out, int32_t out_ld, int32_t flags, void *stream);
/* a6  ME.MinkowskiGenerativeConvolutionTranspose(kernel_size=2, stride=2).forward:
 * out[8*i + k] = in[i] @ W[k] + bias. */
int pcgc_convT_k2s2_fwd(const float *in, int32_t in_ld, int64_t n_in, const float *weight, const float *bias,
                        int32_t cin, int32_t cout, float *out, int32_t out_ld, int32_t flags, void *stream);
/* a5/a6/a7 writing, next to the fp32 output, its pre-split half-precision copy (h2 format, see pcgc_split_h2) for a
 * following h2 k=3 layer, fused into the epilogue.  Returns PCGC_ERR_INVALID for shapes whose kernel cannot pair
 * output channels per lane (use pcgc_split_h2 on the fp32 output then). */
int pcgc_conv_h2out_supported(int32_t kind /* 1: k=1, 2: k=2 s=2, 3: transposed k=2 s=2 */, int32_t cin, int32_t cout);
int pcgc_conv_k1_fwd_h2out(const float *in, int32_t in_ld, int64_t n, const float *weight, const float *bias,
                           int32_t cin, int32_t cout, const float *residual, int32_t res_ld, float *out,
                           int32_t out_ld, uint32_t *out_h2, int32_t out_h2_ld, int32_t flags, int32_t *overflow,
                           void *stream);
int pcgc_conv_k2s2_fwd_h2out(const float *in, int32_t in_ld, const uint64_t *in_keys, const int32_t *child_rows,
                             const int32_t *child_off, int64_t n_parents, const float *weight, const float *bias,
                             int32_t cin, int32_t cout, float *out, int32_t out_ld, uint32_t *out_h2,
                             int32_t out_h2_ld, int32_t flags, int32_t *overflow, void *stream);
int pcgc_convT_k2s2_fwd_h2out(const float *in, int32_t in_ld, int64_t n_in, const float *weight, const float *bias,
                              int32_t cin, int32_t cout, float *out, int32_t out_ld, uint32_t *out_h2,
                              int32_t out_h2_ld, int32_t flags, int32_t *overflow, void *stream);

/* ---- one InceptionResNet block (autoencoder.py:52-57) per call: out = cat(conv0_1(relu(conv0_0(x))),
 * conv1_2(relu(conv1_1(relu(conv1_0(x)))))) + x.  The caller resolves once per layer which kernel serves it (route)
 * and passes that kernel's packed weights; the call issues exactly the launches the per-layer entry points would
 * (same kernels, same order, same results) without a host round trip per layer.  Layers 0..2 = conv0_0, conv0_1,
 * conv1_1 (k=3); w1/b1 = conv1_0, conv1_2 (k=1, reference layout).  x_h2 / out_h2 may be NULL when no layer needs /
 * the caller does not want the h2 copy.  ws: pcgc_irn_ws_bytes(n, c) bytes of device memory for the temporaries. */
enum { PCGC_ROUTE_H2_GATHER = 0, PCGC_ROUTE_H2_OCTET = 1, PCGC_ROUTE_TF32_GATHER = 2, PCGC_ROUTE_TF32_OCTET = 3, PCGC_ROUTE_FP32 = 4,
       PCGC_ROUTE_WIDE = 5 /* tcgen05 / TMA kernel (pcgc_conv_k3_wide_fwd): h2 features, child map */ };
/* PCGC_IRN_MERGED_FIRST: w3[0] / b3[0] / inv_scale[0] describe ONE k=3 convolution c -> c/2 that computes conv0_0 (output
 * channels 0..c/4-1) and conv1_0 (channels c/4..c/2-1: its k=1 weights at the centre offset 13, zeros elsewhere) -- both read x
 * and both are followed by a ReLU (autoencoder.py:52-57); w1[0] / b1[0] are ignored.  All three routes must be h2 routes. */
#define PCGC_IRN_MERGED_FIRST 1
/* PCGC_IRN_FUSED_TAIL: conv1_1 (k=3) and conv1_2 (k=1) run as one kernel (pcgc_conv_k3_octet_h2_k1_fwd); route[2] must be
 * PCGC_ROUTE_H2_OCTET and the shape supported. */
#define PCGC_IRN_FUSED_TAIL 2
/* PCGC_IRN_DUAL_SECOND (with PCGC_IRN_MERGED_FIRST, c = 16, routes 1 and 2 PCGC_ROUTE_H2_OCTET): conv0_1 and conv1_1 + conv1_2
 * run as one kernel (pcgc_irn16_second_stage_fwd): the block is two launches. */
#define PCGC_IRN_DUAL_SECOND 4
typedef struct pcgc_irn_args {
    int64_t n;                      /* rows of the coordinate set */
    int32_t c;                      /* block channels (16, 32, 64) */
    int32_t reserved;               /* flags: PCGC_IRN_MERGED_FIRST */
    const int32_t *nbr;             /* [27][n] kernel map of the set (gather routes), or NULL */
    const int32_t *parent_nbr;      /* [27][n/8] kernel map of the parent set (octet routes), or NULL */
    const float *x;                 /* block input fp32, leading dimension x_ld */
    const uint32_t *x_h2;           /* block input h2, leading dimension x_h2_ld */
    float *out;                     /* block output fp32 [n][c] */
    uint32_t *out_h2;               /* block output h2, or NULL */
    int32_t x_ld, x_h2_ld, out_ld, out_h2_ld;
    int32_t route[3];               /* PCGC_ROUTE_* of conv0_0, conv0_1, conv1_1 */
    float inv_scale[3];             /* of the h2 routes */
    const void *w3[3];              /* packed weights for the route (fp32 `kernel` for PCGC_ROUTE_FP32) */
    const float *b3[3];
    const float *w1[2];
    const float *b1[2];
    void *ws;
    size_t ws_bytes;
    int32_t *overflow;              /* device flag of the h2 kernels, or NULL */
} pcgc_irn_args;
size_t pcgc_irn_ws_bytes(int64_t n, int32_t c);
int pcgc_irn_fwd(const pcgc_irn_args *args, void *stream);

/* ---- backward passes (row a16; MinkowskiEngine Convolution*Backward driven by trainer.py:136) -----
 * Input gradients of the k=3 and k=1 convolutions are forward convolutions of grad_out with the
 * transposed weights (offset-flipped for k=3: W'[k] = W[26-k]^T; stride-1 kernel maps are symmetric),
 * so only the operations without a forward twin are exported here. */

/* Weight and bias gradients are DETERMINISTIC: blocks write partial sums to the caller's workspace
 * (pcgc_conv_bwd_weight_ws_bytes(n, kvol, cin, cout) with n = the row count the pairs are enumerated over, kvol = 8 for
 * the two k=2 layers; pcgc_colsum_ws_bytes()) and a second pass adds them in block order -- no atomics. */
size_t pcgc_conv_bwd_weight_ws_bytes(int64_t n, int32_t kvol, int32_t cin, int32_t cout);
size_t pcgc_colsum_ws_bytes(void);
/* grad_weight [kvol][cin][cout] = sum over pairs of in[a]^T (x) grad_out[b]; kvol 27 with the kernel
 * map of pcgc_kernel_map_k3, or kvol 1 with nbr == NULL (k=1 convolution). */
int pcgc_conv_bwd_weight(const float *in, int32_t in_ld, const int32_t *nbr, int64_t n, int32_t kvol,
                         const float *grad_out, int32_t go_ld, int32_t cin, int32_t cout,
                         float *grad_weight, void *ws, size_t ws_bytes, void *stream);
/* k=2 s=2 down convolution: grad_in[u] = grad_out[parent_of[u]] @ W[key&7]^T, grad_weight [8][cin][cout]
 * (either output may be NULL). */
int pcgc_conv_k2s2_bwd(const float *in, int32_t in_ld, const uint64_t *in_keys, const int32_t *parent_of,
                       int64_t n_in, const float *grad_out, int32_t go_ld, const float *weight, int32_t cin,
                       int32_t cout, float *grad_in, int32_t gi_ld, float *grad_weight, void *ws, size_t ws_bytes,
                       void *stream);
/* generative k=2 s=2 up convolution: grad_in[i] = sum_k grad_out[8i+k] @ W[k]^T, grad_weight [8][cin][cout]. */
int pcgc_convT_k2s2_bwd(const float *in, int32_t in_ld, int64_t n_in, const float *grad_out, int32_t go_ld,
                        const float *weight, int32_t cin, int32_t cout, float *grad_in, int32_t gi_ld,
                        float *grad_weight, void *ws, size_t ws_bytes, void *stream);
/* out[c] = sum over rows of x[r][c]  (bias gradients). */
int pcgc_colsum(const float *x, int32_t ld, int64_t n, int32_t c, float *out, void *ws, size_t ws_bytes, void *stream);

/* ---- selection / pruning (rows a8, a9) --------------------------------------------------- */

/* a9  istopk -- data_utils.py:77-89 (torch.topk on the CPU in the reference): mask[i] = 1 on
 * the k largest logits (stride ld floats apart); ties at the threshold resolve to the lowest rows. */
size_t pcgc_topk_mask_ws_bytes(int64_t n);
int pcgc_topk_mask(const float *logits, int32_t ld, int64_t n, int64_t k, uint8_t *mask, void *ws,
                   size_t ws_bytes, void *stream);
/* a8  ME.MinkowskiPruning()(x, mask) -- autoencoder.py:237,247: stable compaction of keys and
 * feature rows; n_kept (device int32) receives the count.  If nbr_in ([27][n], the k=3 kernel
 * map of the unpruned set) and nbr_out are given, the kernel map of the pruned set is produced
 * too, as [27][n_kept] (compact, caller provides 27*n entries): no re-hashing after pruning. */
size_t pcgc_prune_ws_bytes(int64_t n);
int pcgc_prune(const uint8_t *mask, int64_t n, const uint64_t *keys, const float *feats, int32_t ld,
               int32_t channels, uint64_t *keys_out, float *feats_out, int32_t out_ld, int32_t *n_kept,
               const int32_t *nbr_in, int32_t *nbr_out, void *ws, size_t ws_bytes, void *stream);

/* ---- entropy bottleneck (rows a12, a13, a14) ---------------------------------------------
 * params: the 12 reference tensors entropy_bottleneck._matrices.0..3, _biases.0..3,
 * _factors.0..3 packed per channel as 44 floats (+4 pad = 48):
 *   [M0(3) M1(9, [out][in]) M2(9) M3(3) | B0(3) B1(3) B2(3) B3(1) | F0(3) F1(3) F2(3) F3(1)]
 * (raw values; softplus / tanh are applied inside, entropy_model.py:95-99). */
#define PCGC_EB_PARAMS_PER_CHANNEL 48
/* a12 EntropyBottleneck._likelihood -- entropy_model.py:112-130.  values/lik float32 [n][c]. */
int pcgc_eb_likelihood_fwd(const float *values, int64_t n, int32_t channels, const float *params,
                           float *likelihood, void *stream);
/* a12 backward: grad_values [n][c] (may be NULL) and grad_params [c][48] (same packing as params; raw-parameter
 * gradients, i.e. through softplus / tanh) from grad_likelihood [n][c].  channels must divide 256. */
int pcgc_eb_likelihood_bwd(const float *values, int64_t n, int32_t channels, const float *params,
                           const float *grad_likelihood, float *grad_values, float *grad_params, void *stream);
/* a13/a14 table of compress()/decompress() -- entropy_model.py:155-171,181-189: for symbols
 * min_v..max_v: pmf = max(likelihood, 1e-9); cdf = clamp(cumsum, max=1) with a leading 0
 * -> cdf_float [c][L+1]; cdf_u16 (may be NULL) the torchac integer table (Appendix B.1). */
int pcgc_eb_cdf_table(const float *params, int32_t channels, int32_t min_v, int32_t max_v,
                      float *cdf_float, uint16_t *cdf_u16, void *stream);
/* a13 quantiser of compress() -- entropy_model.py:152-163: minmax (device int32[2], caller
 * initialises to {INT32_MAX, INT32_MIN}) gets min/max of round(feats); second call writes
 * sym = round(feats) - min as int16. */
int pcgc_eb_round_minmax(const float *feats, int64_t count, int32_t *minmax, void *stream);
int pcgc_eb_symbols(const float *feats, int64_t count, const int32_t *minmax, int16_t *sym, void *stream);

/* per-symbol coding intervals on the device (section 8b `pcgc_symbol_ranges`; torchac's per-symbol table walk, Appendix
 * B.2): symbol i is looked up in row (i % n_tables) of the DEVICE uint16 table [n_tables][lp];
 * ranges[i] = c_low | (c_high - 1) << 16 (c_high = 0x10000 for the last symbol).  *bad is set to 1 if a symbol lies outside
 * [0, lp-2] (caller zero-initialises).  pcgc_rc_encode_ranges_host() below codes such a list without touching a table. */
int pcgc_symbol_ranges(const int16_t *sym, int64_t n_sym, const uint16_t *cdf_u16, int32_t n_tables, int32_t lp,
                       uint32_t *ranges, int32_t *bad, void *stream);

/* ---- range coder (row a15; HOST functions, synchronous) ----------------------------------
 * torchac.encode_float_cdf / decode_float_cdf -- entropy_model.py:174,192.
 * cdf_float_host: float32 [n_tables][lp]; symbol i is coded with table row (i % n_tables)
 * (n_tables = channels for the reference's tiled table, = n_sym for per-symbol rows).
 * encode returns the byte count (writes at most cap bytes) or <0. */
int64_t pcgc_rc_encode_host(const float *cdf_float_host, int64_t n_tables, int32_t lp,
                            const int16_t *sym_host, int64_t n_sym, uint8_t *out_host, int64_t cap);
int pcgc_rc_decode_host(const float *cdf_float_host, int64_t n_tables, int32_t lp, const uint8_t *in_host,
                        int64_t in_len, int16_t *sym_host, int64_t n_sym);
/* the same stream from per-symbol intervals (pcgc_symbol_ranges) */
int64_t pcgc_rc_encode_ranges_host(const uint32_t *ranges_host, int64_t n_sym, uint8_t *out_host, int64_t cap);
/* same with a ready uint16 table [n_tables][lp] */
int64_t pcgc_rc_encode_u16_host(const uint16_t *cdf_u16_host, int64_t n_tables, int32_t lp,
                                const int16_t *sym_host, int64_t n_sym, uint8_t *out_host, int64_t cap);
int pcgc_rc_decode_u16_host(const uint16_t *cdf_u16_host, int64_t n_tables, int32_t lp,
                            const uint8_t *in_host, int64_t in_len, int16_t *sym_host, int64_t n_sym);

/* ---- k=3 convolution on the 5th-generation tensor cores (row a3, wide layers; csrc/conv_wide.cuh) ----
 * ME.MinkowskiConvolution(k=3) forward -- autoencoder.py:13,20,35,71,90,109,128,162,174,189,201,216,228 -- for
 * cin in {8,16,32,64}: tcgen05.mma (kind::f16, accumulators in tensor memory), weight tiles streamed
 * by TMA bulk copies, h2 feature rows gathered verbatim into the swizzled K-major operand layout.
 * Same contract as pcgc_conv_k3_h2_fwd (h2 features in, fp32 and/or h2 rows out, fused
 * bias / residual / ReLU, overflow flag); `packed` comes from pcgc_conv_k3_wide_pack_weights
 * (pcgc_conv_k3_wide_packed_bytes bytes, 0 = no kernel for the shape; weights pre-scaled by the
 * power of two `scale`, the epilogue multiplies by inv_scale). */
size_t pcgc_conv_k3_wide_packed_bytes(int32_t cin, int32_t cout);
int pcgc_conv_k3_wide_pack_weights(const float *weight, int32_t cin, int32_t cout, float scale, void *packed, void *stream);
int pcgc_conv_k3_wide_fwd(const uint32_t *feats_h2, int32_t in_ld, const int32_t *nbr, int64_t n, const void *packed,
                          float inv_scale, const float *bias, int32_t cin, int32_t cout, const float *residual,
                          int32_t res_ld, float *out, int32_t out_ld, uint32_t *out_h2, int32_t out_h2_ld, int32_t flags,
                          int32_t *overflow, void *stream);

/* ---- D1 point-to-point distortion (row f3) --------------------------------------------------
 * pc_error(infile1, infile2, res) -- pc_error.py:44-54 (pc_error_d -a A -b B --hausdorff=1
 * --resolution=res-1), read back at coder.py:181-184.  One direction: for every query voxel the
 * squared distance to the nearest voxel of the other cloud (given as its hash table AND its key
 * list), exact.  acc3 (device uint64[3]) receives {sum of squared distances, max squared
 * distance, number of queries that needed the brute-force pass}; open_list: int32 [n_query]
 * scratch.  mse = acc3[0] / n_query; D1 PSNR = 10 log10(3 (res-1)^2 / max(mse_AB, mse_BA)). */
int pcgc_d1_sqdist(const uint64_t *query_keys, int64_t n_query, const uint64_t *table_keys, int64_t cap,
                   const uint64_t *cloud_keys, int64_t n_cloud, int32_t max_radius, uint64_t *acc3, int32_t *open_list,
                   void *stream);

/* ---- k=3 convolution on FULL-OCTET sets on the 5th-generation tensor cores (row a3; csrc/conv_octet_tc05.cuh) ----
 * Same layer and contract as pcgc_conv_k3_octet_h2_fwd (h2 features of the 8-child expansion in, the PARENT set's
 * kernel map, fp32 and/or h2 rows out, fused bias / residual / ReLU, overflow flag) for cin = 16, cout in
 * {1,4,8,16}: the 4x4x4 halos of 32 octets are staged once per tile and the 27 kernel offsets are 27 descriptor
 * start addresses of tcgen05.mma (M = 64) into that halo; weights resident (TMA bulk copy), accumulators in TMEM.
 * `packed` from pcgc_conv_k3_octet_tc05_pack_weights (.._packed_bytes bytes; 0 = no kernel for the shape). */
size_t pcgc_conv_k3_octet_tc05_packed_bytes(int32_t cin, int32_t cout);
int pcgc_conv_k3_octet_tc05_pack_weights(const float *weight, int32_t cin, int32_t cout, float scale, void *packed, void *stream);
int pcgc_conv_k3_octet_tc05_fwd(const uint32_t *feats_h2, int32_t in_ld, const int32_t *parent_nbr, int64_t n_parents,
                                const void *packed, float inv_scale, const float *bias, int32_t cin, int32_t cout,
                                const float *residual, int32_t res_ld, float *out, int32_t out_ld, uint32_t *out_h2,
                                int32_t out_h2_ld, int32_t flags, int32_t *overflow, void *stream);

/* a3 for the FIRST layer (encoder.conv0, autoencoder.py:71,138): the input features are the constant 1 of every voxel
 * (data_utils.py:94, coder.py:131), so out[u] = bias + sum over PRESENT neighbours k of weight[k][0][:] -- computed from the
 * PARENT set's kernel map and child-occupancy info (arguments as pcgc_kernel_map_k3_from_parent, non-full-octet mode) without
 * ever building the 27 x n kernel map of the set.  cout = 16; out fp32 and / or out_h2 (either may be NULL). */
int pcgc_conv_k3_ones_from_parent_fwd(const uint64_t *child_keys, const int32_t *parent_of, const uint64_t *parent_info,
                                      const int32_t *parent_nbr, int64_t n_parents, int64_t n, const float *weight, const float *bias,
                                      int32_t cout, float *out, int32_t out_ld, uint32_t *out_h2, int32_t out_h2_ld, int32_t flags,
                                      int32_t *overflow, void *stream);

/* k=3 convolution + ReLU + the k=1 convolution that follows it, in one kernel (conv1_1 -> conv1_2 of an InceptionResNet
 * block, autoencoder.py:28,42,55): out = relu(conv_k3(in; packed, bias)) @ tail_weight[cmid][cout] + tail_bias + residual.
 * The intermediate is never written.  Shapes: pcgc_conv_k3_octet_h2_k1_supported (4 -> 4 -> 8 today); arguments otherwise
 * as pcgc_conv_k3_octet_h2_fwd. */
int pcgc_conv_k3_octet_h2_k1_supported(int32_t cin, int32_t cmid, int32_t cout);
int pcgc_conv_k3_octet_h2_k1_fwd(const uint32_t *in_h2, int32_t in_ld, const int32_t *parent_nbr, int64_t n_parents,
                                 const uint32_t *packed, float inv_scale, const float *bias, int32_t cin, int32_t cmid,
                                 const float *tail_weight, const float *tail_bias, int32_t cout, const float *residual, int32_t res_ld,
                                 float *out, int32_t out_ld, uint32_t *out_h2, int32_t out_h2_ld, int32_t *overflow, void *stream);

/* Second stage of a 16-channel InceptionResNet block in ONE kernel (autoencoder.py:52-57 with channels = 16):
 *     out[:, 0:8]  = conv0_1(a) + x[:, 0:8]                       (k=3, 4 -> 8)
 *     out[:, 8:16] = conv1_2(relu(conv1_1(b))) + x[:, 8:16]      (k=3, 4 -> 4, then k=1, 4 -> 8)
 * ab_h2 [8 n_parents][>= 8 words]: a | b, the h2 rows the merged first stage wrote (PCGC_IRN_MERGED_FIRST); every 32-byte row is
 * staged once into two CIN = 4 halos.  packed01 / packed11: pcgc_conv_k3_h2_pack_weights of the 4 -> 8 / 4 -> 4 kernels;
 * weight12 fp32 [4][8]; x / out fp32 [n][16] (the block input and output), out_h2 the h2 copy of out (either output may be NULL). */
int pcgc_irn16_second_stage_fwd(const uint32_t *ab_h2, int32_t ab_ld, const int32_t *parent_nbr, int64_t n_parents,
                                const uint32_t *packed01, float inv_scale01, const float *bias01, const uint32_t *packed11,
                                float inv_scale11, const float *bias11, const float *weight12, const float *bias12, const float *x,
                                int32_t x_ld, float *out, int32_t out_ld, uint32_t *out_h2, int32_t out_h2_ld, int32_t *overflow,
                                void *stream);

/* ---- occupancy loss of the training path (row f4) ----------------------------------------------
 * get_bce(data, ground_truth) -- loss.py:7-15 (trainer.py:127-130): isin(data.C, ground_truth.C)
 * fused with BCE-with-logits: for every candidate row the ground-truth table is probed, the
 * stable binary cross entropy is summed (in bits: / ln 2, i.e. the reference's `sum_bce`) and
 * grad_unit[i] = (sigmoid(x_i) - t_i) / ln 2 is written for the backward pass (either output
 * pointer may be NULL; `target` receives the 0/1 mask).  Deterministic two-stage reduction.
 * logits: float [n] with row stride ld; ws: pcgc_bce_isin_ws_bytes() bytes. */
size_t pcgc_bce_isin_ws_bytes(void);
int pcgc_bce_isin(const float *logits, int32_t ld, const uint64_t *cand_keys, int64_t n, const uint64_t *gt_table_keys,
                  int64_t cap, float *loss_sum_bits, float *grad_unit, uint8_t *target, void *ws, size_t ws_bytes,
                  void *stream);

/* ---- ASCII PLY geometry I/O (row f2; HOST functions, synchronous) ---------------------------
 * read_ply_ascii_geo / write_ply_ascii_geo -- data_utils.py:19-48 (coder.py:26,33,128,177).
 * The reader keeps the reference's line semantics: a line is split at single spaces; if any
 * token fails to parse as a float the whole line is skipped (header, comments); the first
 * three values are truncated toward zero.  count_lines gives the row capacity to allocate;
 * parse returns the number of vertex rows written to coords_host (int32 [rows][3]);
 * format writes header + "x y z\n" lines and returns the byte count (cap >= 160 + 36 n). */
int64_t pcgc_ply_count_lines_host(const char *text_host, int64_t len);
int64_t pcgc_ply_parse_ascii_host(const char *text_host, int64_t len, int32_t *coords_host, int64_t cap_rows);
int64_t pcgc_ply_format_ascii_host(const int32_t *coords_host, int64_t n, char *text_host, int64_t cap);

/* ---- coordinate side channel (row f1; HOST functions, synchronous) --------------------------
 * CoordinateCoder.encode / decode -- coder.py:17-36 (gpcc.py:6-36 -> external tmc3 subprocess).
 * In-process lossless octree coder of an int32 [n][3] point SET (non-negative coordinates
 * < 2^21; duplicates collapse; the decoder returns Morton order).  Own stream format ("PCO1"),
 * not G-PCC: see csrc/octree_coder.cpp.  encode returns the byte count (cap >= 9 + 8 n is
 * always enough); decode with coords_host == NULL returns the point count. */
int64_t pcgc_octree_encode_host(const int32_t *coords_host, int64_t n, uint8_t *out_host, int64_t cap);
int64_t pcgc_octree_decode_host(const uint8_t *in_host, int64_t len, int32_t *coords_host, int64_t cap_rows);

#ifdef __cplusplus
}
#endif
#endif /* PCGC_H_ */
