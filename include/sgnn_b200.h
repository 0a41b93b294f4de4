/* sgnn_b200.h -- C ABI of libsgnn_b200.so (hand-written sm_100a kernels).
 *
 * Drop-in boundary for the sparse-convolution + generative-upsampling hot path that
 * SG-NN's torch/model.py reaches through `import sparseconvnet as scn` (reference
 * torch/model.py:7).  The reference binds that path through scn's pybind module
 * (upstream sparseconvnet/SCN/pybind.cpp -- NOT present under /root/reference); each
 * entry point below names the reference call site(s) in torch/model.py it serves and
 * the upstream scn function whose role it takes.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every pointer marked "dev" is a CUDA device pointer owned by the caller;
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it
 *     (no host synchronisation, no allocation, no mutable global state: tuning lives in argument structs) and is re-entrant for
 *     distinct (buffers, stream) pairs.  The one exception is sgnn_generator_forward, which orchestrates a whole pass: it reads
 *     nine data-dependent row counts back (4 bytes each, pinned), keeps per-thread caches (pinned slots, events, a side stream
 *     per device) and, for small levels, forks work to that side stream and joins it back before it returns;
 *   - return value: SGNN_OK or a negative SGNN_E_* code; nothing throws across the ABI;
 *   - features are row-major [n_rows, ld] with `ld` (elements) >= channels;
 *   - coordinates are int32 [n,4] = (z, y, x, batch)  (model.py:321, scene_dataloader.py:17,30);
 *   - filter offsets are numbered row-major over (dz,dy,dx), last fastest:
 *       3^3 submanifold: k = (dz+1)*9 + (dy+1)*3 + (dx+1);   2^3 stride 2: k = (z&1)*4 + (y&1)*2 + (x&1);
 *   - weights are [K, Cin, Cout] exactly as scn stores them (App. A.4);
 *   - rulebooks are OUTPUT-STATIONARY neighbour tables nbr[k][row] (k-major, -1 = absent),
 *     the transpose of scn's per-offset (in,out) pair lists: pair list k == {(nbr[k][j], j) : nbr[k][j] >= 0}.
 */
#ifndef SGNN_B200_H
#define SGNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGNN_VERSION 100

enum {
  SGNN_OK = 0,
  SGNN_E_INVALID = -1,     /* bad argument (null pointer, negative size, unsupported channel count) */
  SGNN_E_CUDA = -2,        /* a CUDA runtime call failed; sgnn_last_cuda_error() has the code */
  SGNN_E_TOO_LARGE = -3,   /* extent or row count exceeds the 2^31-1 indexing of this build */
  SGNN_E_UNSUPPORTED = -4, /* valid in scn, not implemented here (e.g. filter size other than 3 / 2) */
  SGNN_E_ALIGN = -5,       /* pointer / leading dimension breaks an alignment rule stated below */
  SGNN_E_NOMEM = -6        /* caller-provided workspace arena too small */
};

enum { SGNN_F32 = 0, SGNN_BF16 = 1 };

/* Conv tile height: rows per CTA of sgnn_conv_forward. */
#define SGNN_CONV_TILE 128

/* Active-site set at one resolution: a bitmask over the (batch, z, y, x) extent, 64 x-cells per
 * word, plus an exclusive popcount prefix (raster rank) and an optional rank->row permutation.
 * Plays the role of scn's Metadata SparseGrid hash maps.  The struct lives on
 * the host; the three arrays are device memory supplied by the caller:
 *   mask        [n_words]      uint64
 *   prefix      [n_words + 1]  int32   prefix[n_words] == number of active cells
 *   row_of_rank [n_rows]       int32   or NULL when row id == raster rank                        */
typedef struct SgnnGrid {
  int32_t nb, d0, d1, d2;  /* batch count and extent in cells along z, y, x */
  int32_t wx;              /* words per x-row = (d2 + 63) / 64 */
  int32_t reserved;
  int64_t n_words;         /* nb * d0 * d1 * wx */
  uint64_t* mask;
  int32_t* prefix;
  int32_t* row_of_rank;
} SgnnGrid;

/* Bytes of scratch the grid/scan/compaction calls need for `n_items` scanned items. */
size_t sgnn_scan_scratch_bytes(int64_t n_items);
/* Scratch for the compaction calls (sgnn_dense_to_sparse, sgnn_heads_compact) over n_items candidates. */
size_t sgnn_compact_scratch_bytes(int64_t n_items);

/* ---- a1: scn.InputLayer(3, size, mode=0)  (model.py:51,216,224,265; upstream InputLayer_updateOutput)
 * Builds mask/prefix/row_of_rank from caller-ordered coordinates.  coords: dev [n,4], int64 when
 * coords_i64 != 0 (the LongTensor the reference passes) else int32.  coords_i32_out: dev [n,4] int32 copy
 * or NULL.  A coordinate outside the extent sets *status (dev int32, may be NULL) to 1 and is ignored.
 * Duplicate coordinates: the highest row id owns the cell (scn: later insert overwrites, App. A.2). */
int sgnn_grid_build(const SgnnGrid* g, const void* coords, int coords_i64, int64_t n,
                    int32_t* coords_i32_out, int32_t* status, void* scratch, size_t scratch_bytes,
                    void* stream);

/* ---- a4 (site set): scn.Convolution(3,C,C,2,2) output sites (model.py:44; upstream
 * Convolution_InputSgsToRulesAndOutputSgs).  coarse cell = fine cell >> 1; coarse rows are numbered in
 * raster order (upstream order is implementation defined), so coarse->row_of_rank is
 * unused.  coarse extent <= ceil(fine extent / 2) per axis; fine cells whose parent falls outside the coarse
 * extent are dropped, as scn drops parents >= its (S-2)/2+1 output size. */
int sgnn_grid_coarsen(const SgnnGrid* fine, const SgnnGrid* coarse, void* scratch,
                      size_t scratch_bytes, void* stream);

/* ---- a7: metadata.getSpatialLocations (model.py:380).  Coordinates of a raster-ordered grid. */
int sgnn_grid_enumerate(const SgnnGrid* g, int32_t* coords_out, void* stream);

/* Row lookup: rows_out[i] = row of (coords[i].zyx >> shift, batch) or -1. */
int sgnn_grid_lookup(const SgnnGrid* g, const int32_t* coords, int64_t n, int shift,
                     int32_t* rows_out, void* stream);

/* ---- a2: submanifold 3^3 rulebook (first SubmanifoldConvolution per resolution, model.py:53,217,225,266;
 * upstream Metadata::getSubmanifoldRuleBook).  nbr: dev [27][n] int32. */
int sgnn_rulebook_submanifold(const SgnnGrid* g, const int32_t* coords, int64_t n, int32_t* nbr,
                              void* stream);
/* The same rulebook in COMPACT form, for site sets whose rows have few neighbours (the 5 %-occupancy encoder input,
 * model.py:32,38,40: 2.3 of 27 offsets present -- the dense table is 108 B/site of which 9 B are rules): per site only the
 * PRESENT offsets, in ascending k (centre included), packed k << 27 | input_row:
 *   cnt   dev [n] uint8      number of present offsets of the site (1..27)
 *   slots dev [27][n] int32  slots[s][i], s < cnt[i]; entries at s >= cnt[i] are NOT written (allocate 27 * n, touch few)
 * upstream: the same Metadata::getSubmanifoldRuleBook, regrouped per output row.  n < 2^27. */
int sgnn_rulebook_submanifold_compact(const SgnnGrid* g, const int32_t* coords, int64_t n, int32_t* slots, uint8_t* cnt,
                                      void* stream);

/* ---- a4 (rules): filter 2 stride 2 rulebook (upstream Metadata::getRuleBook).
 *   parent   dev [n_fine]        coarse_row * 8 + k, or -1 if the fine site has no coarse parent
 *   children dev [8][n_coarse]   fine row at offset k of coarse row, or -1                       */
int sgnn_rulebook_strided(const SgnnGrid* coarse, const int32_t* fine_coords, int64_t n_fine,
                          int32_t* parent, int32_t* children, int64_t n_coarse, void* stream);
/* sgnn_grid_enumerate(coarse) + sgnn_rulebook_strided(coarse, fine) in one kernel, from the coarse side (no memset of the
 * children table; `parent` is preset to -1 only when an odd fine extent leaves fine sites without a parent).  Requires a
 * raster-ordered coarse set (sgnn_grid_coarsen) and a fine set without duplicate coordinates (duplicate rows of a mode-0 input
 * keep the parent of the row that owns the cell only -- use the two calls there).  Same outputs otherwise. */
int sgnn_grid_coarse_build(const SgnnGrid* fine, const SgnnGrid* coarse, int64_t n_fine, int64_t n_coarse,
                           int32_t* coarse_coords, int32_t* parent, int32_t* children, void* stream);

/* Epilogue slot of a convolution: y = acc (+ residual); if scale: y = y*scale[c] + shift[c]
 * (one fused multiply-add); if relu: y = max(y, 0).  out == NULL disables the slot. */
typedef struct SgnnEpilogue {
  void* out;           /* dev [n_out, ld] */
  int32_t ld;
  int32_t relu;
  const float* scale;  /* dev [cout] or NULL */
  const float* shift;  /* dev [cout] or NULL */
} SgnnEpilogue;

/* ---- a3 / a4 / a9: sparse convolution forward, gather -> K x (Cin x Cout) contraction, no scatter
 * (model.py:32,38,40,44,179,186,254; upstream SubmanifoldConvolution_updateOutput / Convolution_updateOutput).
 *   out[j] = sum_{k ascending} in[nbr[k][j] >> gather_shift] @ W[k]    over nbr[k][j] >= 0
 * fp32 path: accumulation is one fmaf chain per output element, k ascending then ci ascending, from +0
 * (bit-reproducible; oracle/o3.c restates it).  Cout in {4,8,12,16}; Cin <= 64.
 * child_mode = 1 evaluates the convolution on the 8 children of every row of a parent set straight from the
 * PARENT rulebook (generative upsampling, model.py:192-207,224-225): output row 8*p+c (c = z-major child
 * index), offset (dz,dy,dx) of child c reads parent neighbour floor((c_a+d_a)/2) per axis; the x8 replicated
 * features of model.py:202 never exist in memory.
 * Alignment: when ld_in % 4 == 0 and `in` is 16-byte aligned the gather moves 16-byte chunks, otherwise
 * 4-byte elements; outputs/residual need ld % 4 == 0 and 16-byte alignment (else SGNN_E_ALIGN). */
typedef struct SgnnConvArgs {
  const void* in;         /* dev [n_in, ld_in] */
  int32_t ld_in;
  int32_t dtype;          /* SGNN_F32 | SGNN_BF16 (features and weights) */
  const int32_t* nbr;     /* dev [K][nbr_stride] */
  int64_t nbr_stride;
  int32_t K;              /* 27 or 8 */
  int32_t child_mode;     /* 0: plain; 1: outputs are the 8 children of each nbr row (K must be 27) */
  const void* weight;     /* dev [K][cin][cout] */
  int32_t cin, cout;
  int64_t n_out;          /* output rows (8 * parent rows in child mode) */
  const void* residual;   /* dev [n_out, ld_res] or NULL */
  int32_t ld_res;
  int32_t n_in;           /* rows of `in`, or 0 if not stated (informational) */
  SgnnEpilogue a, b;
  int32_t flags;          /* SGNN_CONV_* bits */
  int32_t reserved;
} SgnnConvArgs;
/* tensor-core calls: `workspace` already holds this weight's prepared filter bank (sgnn_conv_tc32_prepare): skip the preparation */
#define SGNN_CONV_PREPARED 1
/* sgnn_conv_forward kernel selection (A/B; the result bits do not depend on it): force / forbid the lane = (row, channel
 * group) kernel of csrc/conv_sp.cu (K = 27 only), by default used for K = 27 launches of <= 4096 rows */
#define SGNN_CONV_ROWLANE 2
#define SGNN_CONV_NO_ROWLANE 4
int sgnn_conv_forward(const SgnnConvArgs* args, void* stream);
/* sgnn_conv_forward (fp32, K = 27, not child mode) over a compact rulebook (sgnn_rulebook_submanifold_compact): args->nbr is
 * ignored, args->nbr_stride is the slot stride (n).  Visits only the present offsets of a row; the same fmaf chain (k ascending,
 * ci ascending, absent offsets skipped as oracle/o3.c skips them) => bit-identical to sgnn_conv_forward.  (Cout, Cin) in
 * {(8,1), (8,8), (12,8), (12,12), (16,12), (16,16)}, else SGNN_E_UNSUPPORTED.  csrc/conv_sp.cu. */
int sgnn_conv_forward_compact(const SgnnConvArgs* args, const int32_t* slots, const uint8_t* cnt, void* stream);

/* ---- a3 / a4 / a9 on the tensor cores: the same operation as sgnn_conv_forward for fp32 features with Cout = 16
 * and Cin <= 48 (every wide layer of the generator: model.py:179,186,254 and the FullyConvolutionalNet blocks),
 * evaluated with tcgen05.mma on an exact 3-way bf16 split of features and filters (six bf16 partial products per
 * fp32 product, fp32 accumulation in TMEM; csrc/conv_tc32.cu).  fp32 accuracy (relative ~1e-6 of the row magnitude),
 * but NOT the fixed fmaf order of sgnn_conv_forward -- results are not bit-identical to it.  In child mode the 27
 * filter offsets of each child collapse onto its 8 parent neighbours (filters pre-summed per child and parent
 * offset).  `workspace`: dev scratch of sgnn_conv_tc32_workspace_bytes(K, cin, child_mode) bytes for the prepared
 * filter bank (written by the call, on `stream`).  SGNN_E_UNSUPPORTED for shapes outside the above. */
size_t sgnn_conv_tc32_workspace_bytes(int32_t K, int32_t cin, int32_t child_mode);
int sgnn_conv_forward_tc32(const SgnnConvArgs* args, void* workspace, size_t workspace_bytes, void* stream);
/* The filter-bank preparation of the two tensor-core calls on its own (fp32 [K][cin][cout] -> split bf16 planes in the
 * tensor core's operand layout; child mode: pre-summed per child and parent offset).  Weights do not change between
 * forward passes, so a caller may prepare once and pass SGNN_CONV_PREPARED afterwards. */
int sgnn_conv_tc32_prepare(const void* weight, int32_t K, int32_t cin, int32_t cout, int32_t child_mode, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---- a2 + a3, unique-row form (csrc/conv_ur.cu).  A TILE PLAN re-expresses the submanifold rulebook of a site set per
 * 128-row tile: the sorted list of DISTINCT input rows the tile's 27 filter offsets touch and a 27 x 128 table of 16-bit
 * indices into that list (0xFFFF = absent) -- scn's per-offset (in,out) pair lists (upstream
 * Metadata::getSubmanifoldRuleBook) regrouped so that a convolution stages every input row of a tile ONCE.  Built once per
 * site set from its neighbour table, shared by every SubmanifoldConvolution on that set (model.py:38,40,179,186,254 ...).
 * plan: dev, 256-byte aligned, sgnn_tile_plan_bytes(n_rows) bytes, opaque. */
size_t sgnn_tile_plan_bytes(int64_t n_rows);
int sgnn_tile_plan_build(const int32_t* nbr, int64_t nbr_stride, int64_t n_rows, void* plan, size_t plan_bytes,
                         void* stream);
/* sgnn_rulebook_submanifold + sgnn_tile_plan_build in one kernel (the table is written once and not read back): nbr dev
 * [27][n] and the plan of that table.  Same outputs as the two calls. */
int sgnn_rulebook_submanifold_plan(const SgnnGrid* g, const int32_t* coords, int64_t n, int32_t* nbr, void* plan,
                                   size_t plan_bytes, void* stream);
/* sgnn_conv_forward_tc32 for K = 27, Cin <= 32 (ld_in % 4 == 0), Cout in {8, 12, 16} with a tile plan of args->nbr: the distinct rows of a tile are
 * fetched by TMA (cp.async.bulk, one per row) into shared memory, split into the bf16 planes once, and expanded filter
 * offset by filter offset from shared memory into tensor memory for tcgen05.mma.  Same arithmetic and tolerance as
 * sgnn_conv_forward_tc32; `workspace` as there (sgnn_conv_tc32_workspace_bytes(27, cin, 0)). */
int sgnn_conv_forward_tc32_ur(const SgnnConvArgs* args, const void* plan, void* workspace, size_t workspace_bytes,
                              void* stream);

/* Child mode (generative upsampling, model.py:192-207,224-225) of the unique-row form: args->child_mode = 1, Cin = 48, Cout = 16,
 * `plan` = the tile plan of the PARENT site set (args->nbr).  Distinct parent rows of a 128-parent tile are staged once by TMA,
 * the pre-summed child filters stream through a shared-memory ring, children that are consecutive in the accumulator share
 * one tcgen05.mma.  `workspace`: 64 * 4608 bytes; sgnn_conv_urc_prepare fills it once per weight (then SGNN_CONV_PREPARED). */
int sgnn_conv_urc_prepare(const void* weight, int32_t cin, void* workspace, size_t workspace_bytes, void* stream);
int sgnn_conv_forward_tc32_urc(const SgnnConvArgs* args, const void* plan, void* workspace, size_t workspace_bytes,
                               void* stream);

/* ---- scn.Deconvolution(3,Cin,Cout,2,2) (north_star operator surface; upstream Deconvolution_updateOutput)
 *   out[i] = in[parent[i] >> 3] @ W[parent[i] & 7]          (rows with parent < 0 get zeros) */
int sgnn_deconv_forward(const void* in, int32_t ld_in, int32_t dtype, const int32_t* parent,
                        const void* weight, int32_t cin, int32_t cout, int64_t n_fine,
                        const SgnnEpilogue* ep, void* stream);

/* ---- a6: scn.UnPooling(3,2,2) (inside FullyConvolutionalNet, model.py:180,255; upstream UnPooling_updateOutput)
 *   out[i][0:c] = in[parent[i] >> 3][0:c], optional per-channel affine + relu (folds the trailing BatchNormReLU) */
int sgnn_unpool(const float* in, int32_t ld_in, const int32_t* parent, int32_t c, int64_t n_fine,
                const SgnnEpilogue* ep, void* stream);

/* ---- a5: scn.BatchNormReLU eval mode (model.py:37,39,42,45,181,187,256; upstream BatchNormalization_updateOutput)
 *   y = max(fma(x, scale[c], shift[c]), 0)   scale = gamma * rsqrt(var + 1e-4), shift = beta - mean*scale
 * (host folds; relu optional).  x and y may alias. */
int sgnn_affine_relu(const float* x, int32_t ld_x, float* y, int32_t ld_y, int64_t n, int32_t c,
                     const float* scale, const float* shift, int32_t relu, void* stream);

/* ---- a6: AddTable / JoinTable helpers (row-wise add; column-slot copy) */
int sgnn_add_rows(const float* a, int32_t ld_a, const float* b, int32_t ld_b, float* y,
                  int32_t ld_y, int64_t n, int32_t c, void* stream);
int sgnn_copy_cols(const float* src, int32_t ld_src, float* dst, int32_t ld_dst, int64_t n,
                   int32_t c, void* stream);

/* nn.Linear heads (model.py:189-190,230-231,258,271): y[i][o] = (sum_c ascending fma(x[i][c], w[o][c])) + b[o] */
int sgnn_linear(const float* x, int32_t ld_x, const float* w, const float* b, float* y,
                int32_t ld_y, int64_t n, int32_t cin, int32_t cout, void* stream);

/* ---- a7: scn.SparseToDense (model.py:47,65; upstream SparseToDense_updateOutput).
 * dense: dev [nb, c, d0, d1, d2] fp32, fully overwritten (zeros + scatter). */
int sgnn_sparse_to_dense(const float* feats, int32_t ld, const int32_t* coords, int64_t n,
                         int32_t c, float* dense, int32_t nb, int32_t d0, int32_t d1, int32_t d2,
                         void* stream);

/* ---- a12: dense coarse U-Net layers (model.py:89-136,152-166: nn.Conv3d / nn.ConvTranspose3d + BatchNorm3d + ReLU).
 * NCDHW fp32; the input is the channel concatenation of in0 [nb,c0,d0,d1,d2] and (optional) in1 [nb,c1,...]
 * (torch.cat of model.py:156,160).  conv3d weight [cout][c0+c1][k][k][k]; convT3d weight [c0+c1][cout][k][k][k]
 * (PyTorch layouts).  out [nb,cout,o0,o1,o2] with the usual output sizes.  Fixed order: ci ascending, then
 * kz,ky,kx ascending, one fmaf chain; optional folded BatchNorm (scale, shift per cout) and relu. */
int sgnn_dense_conv3d(const float* in0, int32_t c0, const float* in1, int32_t c1, int32_t nb, int32_t d0,
                      int32_t d1, int32_t d2, const float* w, int32_t cout, int32_t ksize, int32_t stride,
                      int32_t pad, const float* scale, const float* shift, int32_t relu, float* out, void* stream);
int sgnn_dense_convT3d(const float* in0, int32_t c0, const float* in1, int32_t c1, int32_t nb, int32_t d0,
                       int32_t d1, int32_t d2, const float* w, int32_t cout, int32_t ksize, int32_t stride,
                       int32_t pad, const float* scale, const float* shift, int32_t relu, float* out, void* stream);

/* ---- a8: GenModel.dense_coarse_to_sparse (model.py:315-336).
 * dense_feats dev [nb,c,d0,d1,d2]; dense_out dev [nb,2,d0,d1,d2] (occ logit, sdf).
 * keep = sigmoid(occ) > 0.5 evaluated literally in fp32.  Kept cells are compacted in raster order:
 *   locs dev [cap,4] int32, feats dev [cap, ld_feats] = [occ, sdf, feats(c)], *count dev int32.
 *   cand_out dev [nb*d0*d1*d2, 2] = (occ, sdf) of every cell (model.py:336), or NULL. */
int sgnn_dense_to_sparse(const float* dense_feats, const float* dense_out, int32_t nb, int32_t c,
                         int32_t d0, int32_t d1, int32_t d2, int32_t* locs, float* feats,
                         int32_t ld_feats, float* cand_out, int32_t* count, void* scratch,
                         size_t scratch_bytes, void* stream);

/* ---- a9: Refinement heads + mask + compaction (model.py:230-247) on the 8*n_parent candidates.
 *   x dev [n_cand, ld_x] post-BatchNormReLU features (16 ch); w_occ/w_sdf dev [c]; b_occ/b_sdf dev [1].
 *   cand_out dev [n_cand,2] = (occ, sdf) for every candidate (model.py:240,247).
 *   kept (sigmoid(occ) > 0.5), candidate order preserved:
 *     locs dev [cap,4] = 2*parent_coords + child offset (z-major child order, model.py:195-198)
 *     feats dev [cap, ld_feats] = [x(c), occ, sdf]   (model.py:242);  *count dev int32 */
int sgnn_heads_compact(const float* x, int32_t ld_x, int32_t c, const float* w_occ,
                       const float* b_occ, const float* w_sdf, const float* b_sdf,
                       const int32_t* parent_coords, int64_t n_parent, float* cand_out,
                       int32_t* locs, float* feats, int32_t ld_feats, int32_t* count,
                       void* scratch, size_t scratch_bytes, void* stream);

/* Two-phase forms of the two compactions: phase 1 computes (occ, sdf), the literal sigmoid mask and its exclusive
 * scan -- offs[n] is the kept count, readable with one 4-byte copy -- phase 2 writes exactly that many rows.
 * flags dev [n] uint8, offs dev [n+1] int32, scratch >= sgnn_scan_scratch_bytes(n). */
int sgnn_heads_flags(const float* x, int32_t ld_x, int32_t c, const float* w_occ, const float* b_occ,
                     const float* w_sdf, const float* b_sdf, int64_t n_cand, float* cand_out, uint8_t* flags,
                     int32_t* offs, void* scratch, size_t scratch_bytes, void* stream);
int sgnn_heads_write(const float* x, int32_t ld_x, int32_t c, const float* cand_out,
                     const int32_t* parent_coords, int64_t n_cand, const uint8_t* flags, const int32_t* offs,
                     int32_t* locs, float* feats, int32_t ld_feats, void* stream);
/* sgnn_heads_write + the skip join of the NEXT level in one pass (sgnn_concat_skip, model.py:338-355,391,401): columns
 * [c+2, c+2+c_skip) of a kept row receive the features of the skip site at the row's own (child) coordinates, zeros where the
 * skip set has none; ld_feats >= c + 2 + c_skip.  Same values as the two calls. */
int sgnn_heads_write_join(const float* x, int32_t ld_x, int32_t c, const float* cand_out, const int32_t* parent_coords,
                          int64_t n_cand, const uint8_t* flags, const int32_t* offs, int32_t* locs, float* feats,
                          int32_t ld_feats, const SgnnGrid* skip_grid, const float* skip_feats, int32_t ld_skip,
                          int32_t c_skip, void* stream);
int sgnn_dense_flags(const float* dense_out, int32_t nb, int64_t vol, float* cand_out, uint8_t* flags,
                     int32_t* offs, void* scratch, size_t scratch_bytes, void* stream);
int sgnn_dense_write(const float* dense_feats, const float* dense_out, int32_t nb, int32_t c, int32_t d0,
                     int32_t d1, int32_t d2, const uint8_t* flags, const int32_t* offs, int32_t* locs,
                     float* feats, int32_t ld_feats, void* stream);

/* ---- the whole generator in one native call: GenModel.forward of model.py:371-416 with the default SG-NN
 * structure (test_scene.py:29-39: 3 sparse encoder levels, dense U-Net, 3 refinement levels, surface head,
 * pass_occ + pass_feats, sparse + dense skips).  Host orchestration (site sets, rulebooks, fused convolutions,
 * dense U-Net, generative upsampling) runs in C++ on `stream`; the only host<->device traffic is one 4-byte read
 * per data-dependent row count.  All weights are device pointers in the layouts of the individual calls above;
 * BatchNorm layers are passed folded (scale, shift).  Everything the call produces lives in the caller's `arena`
 * (device memory); on SGNN_E_NOMEM out->arena_needed is a size that suffices for a retry. */
typedef struct SgnnBnFold { const float* scale; const float* shift; } SgnnBnFold;
typedef struct SgnnResBlockW {      /* ConcatTable(Identity, Seq(BNReLU, SMC, BNReLU, SMC)) + AddTable */
  SgnnBnFold bn0; const float* w0; SgnnBnFold bn1; const float* w1;
} SgnnResBlockW;
typedef struct SgnnEncLevelW {      /* SparseEncoderLayer, model.py:21-67 */
  int32_t cin, c;
  const float* w_in; SgnnResBlockW res; SgnnBnFold bn_out; const float* w_down; SgnnBnFold bn_down;
} SgnnEncLevelW;
typedef struct SgnnFcnW {           /* FullyConvolutionalNet(reps 1, [c,c,c], residual) + BatchNormReLU(3c) */
  int32_t c, reserved;
  SgnnResBlockW blk[3]; SgnnBnFold bn_down[2]; const float* w_down[2]; SgnnBnFold bn_join;
} SgnnFcnW;
typedef struct SgnnDenseLayerW {    /* Conv3d / ConvTranspose3d + BatchNorm3d + ReLU, model.py:89-129 */
  const float* w; SgnnBnFold bn; int32_t cout, ksize, stride, pad, transposed, cat_with; /* cat_with: -1 or layer idx */
} SgnnDenseLayerW;
typedef struct SgnnRefineW {        /* Refinement, model.py:169-247 */
  int32_t cin, c;
  const float* w_in; SgnnFcnW fcn; const float* w_up; SgnnBnFold bn_up;
  const float* w_occ; const float* b_occ; const float* w_sdf; const float* b_sdf;
} SgnnRefineW;
typedef struct SgnnSurfaceW {       /* SurfacePrediction, model.py:249-272 */
  int32_t cin, c;
  const float* w_in; SgnnFcnW fcn; const float* w_lin; const float* b_lin;
} SgnnSurfaceW;
typedef struct SgnnGeneratorW {
  SgnnEncLevelW enc[3];
  SgnnDenseLayerW dense[6];         /* encode0, encode1, bottleneck, decode3 (cat enc1), decode4 (cat enc0), final */
  const float* w_heads;             /* [2][nf_coarse]: occpred, sdfpred (1x1x1, no bias) */
  int32_t nf_coarse, reserved;
  SgnnRefineW ref[3];
  SgnnSurfaceW surf;
  void* prepared;                   /* dev, sgnn_generator_prepared_bytes() bytes, filled by sgnn_generator_prepare; or NULL */
  size_t prepared_bytes;
  /* SGNN_GEN_TC32 row thresholds, 0 = default: site sets with >= ur_min_rows rows (1000) get a unique-row tile plan; other
   * Cout = 16 convolutions with >= tc32_min_rows output rows (60000) use sgnn_conv_forward_tc32, the rest sgnn_conv_forward */
  int64_t tc32_min_rows, ur_min_rows;
  /* levels of at most overlap_max_rows rows build the coarse site sets of their FullyConvolutionalNet on a second stream under
   * their first convolutions (generator.cu fork_side); 0 = default (200000), < 0 = never */
  int64_t overlap_max_rows;
} SgnnGeneratorW;
typedef struct SgnnGeneratorOut {
  int64_t n_out;        int32_t* out_locs;  float* out_sdf;        /* [n_out,4], [n_out,1] */
  int64_t n_cand[4];    int32_t* cand_locs[4]; float* cand[4];     /* per level: [n,4] (if requested), [n,2] */
  int64_t rows[16];     /* site counts: enc L0..L3, then per refinement/surface level its 3 FCN resolutions */
  size_t arena_used, arena_needed;
  double conv_ms;       /* SGNN_GEN_PROFILE: sum of CUDA-event durations of the sgnn_conv_forward launches */
  int64_t n_conv;
} SgnnGeneratorOut;
#define SGNN_GEN_CAND_LOCS 1        /* materialise the candidate coordinates of every level (model.py:247,336) */
#define SGNN_GEN_CAND_PARENTS 32    /* instead: cand_locs[h] = the level's PARENT coordinates [n_cand/8, 4] (for SGNN_EXPORT_CHILDREN) */
#define SGNN_GEN_PROFILE 2          /* time every convolution launch with CUDA events (adds one sync at the end) */
#define SGNN_GEN_TC32 4             /* run the Cout = 16 convolutions through sgnn_conv_forward_tc32 (tensor cores) */
#define SGNN_GEN_PHASES 16          /* one CUDA event per phase boundary of the pass: in-situ time of every phase, launch gaps and
                                     * host reads included (sgnn_generator_phase_entry; adds one sync at the end) */
#define SGNN_GEN_DENSE_RULES 8      /* A/B: dense neighbour table + sgnn_conv_forward on the encoder's input level instead of the
                                     * compact rulebook + sgnn_conv_forward_compact (same bits either way) */
/* Prepared tensor-core filter banks of every Cout = 16 convolution of the generator, built once per weight set: with
 * w->prepared set, SGNN_GEN_TC32 passes launch no preparation kernels (~30 launches per pass otherwise). */
size_t sgnn_generator_prepared_bytes(const SgnnGeneratorW* w);
int sgnn_generator_prepare(const SgnnGeneratorW* w, void* stream);
int sgnn_generator_forward(const SgnnGeneratorW* w, const void* coords, int coords_i64, const float* feats,
                           int64_t n, int32_t nb, const int32_t* dims3, void* arena, size_t arena_bytes, int flags,
                           SgnnGeneratorOut* out, void* stream);

/* Per-convolution record of the calling thread's last SGNN_GEN_PROFILE pass (i = 0 .. n_conv-1, launch order):
 * rec6 = {n_out, cin, cout, K, child_mode, ran on tensor cores}, *ms = CUDA-event duration.  SGNN_E_INVALID past the end. */
int sgnn_generator_profile_entry(int32_t i, int64_t* rec6, float* ms);
/* Phase i of the calling thread's last SGNN_GEN_PHASES pass (i = 1 .. n-1): name48 = label (<= 47 chars + NUL), *ms = CUDA-event
 * time since the previous phase boundary.  SGNN_E_INVALID past the end. */
int sgnn_generator_phase_entry(int32_t i, char* name48, float* ms);

/* ---- marching cubes on the predicted dense TSDF, the step after the forward pass
 * (reference torch/marching_cubes/marching_cubes.cpp:459-517 run_marching_cubes, called from data_util.py:270-284).
 * tsdf dev [n0,n1,n2] fp32 (z,y,x), -inf = unobserved.  Three calls:
 *   sgnn_mc_count   per cell the number of triangles the reference emits for it (8 trilinear corner samples, cube index,
 *                   the reference's threshold tests and table), exclusive-scanned: offs dev [n0*n1*n2 + 1]; offs[last]
 *                   = number of triangles.  scratch >= sgnn_mc_scratch_bytes(n0, n1, n2).
 *   sgnn_mc_emit    the triangle soup, tris dev [n_tri, 3, 3] (x,y,z per vertex), in the reference's order (cells z,y,x
 *                   raster, table order) and with its bits (every float operation rounds once, in its operand order).
 *   sgnn_mc_merge_host  HOST function on host arrays: merge_close_vertices(1e-5, approx) in first-come order, then
 *                   remove_degenerate_faces / remove_duplicate_faces (:266-455).  verts [3*n_tri,3] and faces [n_tri,3]
 *                   are caller-allocated upper bounds; the used counts are returned.
 *   sgnn_mc_merge_host_src  the same, and vert_src[v] (host [3*n_tri], may be NULL) = index into the soup's 3*n_tri
 *                   vertices of the first-come vertex that became merged vertex v -- the reference gives v that vertex's
 *                   colour (:399-431).
 *   sgnn_mc_tri_cells   tri_cell dev [n_tri] = linear (z,y,x) cell index each triangle of the soup was emitted for; the
 *                   reference paints a triangle with the colour of that cell's voxel (:228-231,255-257), so per-voxel
 *                   colours need no kernel of their own: colour[v] = colors[tri_cell[vert_src[v] / 3]]. */
size_t sgnn_mc_scratch_bytes(int32_t n0, int32_t n1, int32_t n2);
int sgnn_mc_count(const float* tsdf, int32_t n0, int32_t n1, int32_t n2, float isovalue, float truncation, float thresh,
                  int32_t* offs, void* scratch, size_t scratch_bytes, void* stream);
int sgnn_mc_emit(const float* tsdf, int32_t n0, int32_t n1, int32_t n2, float isovalue, float truncation, float thresh,
                 const int32_t* offs, float* tris, void* stream);
int sgnn_mc_merge_host(const float* tris, int64_t n_tri, float* verts, int32_t* faces, int64_t* n_verts,
                       int64_t* n_faces);
int sgnn_mc_merge_host_src(const float* tris, int64_t n_tri, float* verts, int32_t* faces, int32_t* vert_src,
                           int64_t* n_verts, int64_t* n_faces);
int sgnn_mc_tri_cells(const int32_t* offs, int64_t n_cells, int32_t* tri_cell, void* stream);
/* The packed triangulation table the kernels use: 256 words, 4 bits per triangle-vertex edge id, 0xF terminates. */
int sgnn_mc_table(uint64_t* out256);

/* ---- result export: every output tensor of a pass in ONE launch.  The generator's results live in its arena; the reference
 * returns fresh tensors with int64 coordinates (model.py:247,336,380,416).  A segment = n units of one kind:
 *   COPY32       n 4-byte elements src -> dst
 *   COORDS       n rows int32 [n,4] -> int32 / int64 (to_i64) [n,4]
 *   CHILDREN     n = 8 * parents: row i = child (i & 7) of parent row i >> 3 of src int32 [n/8,4] (model.py:192-207)
 *   DENSE_CELLS  n = nb*d0*d1*d2 cells of a dense grid, batch-major raster order (model.py:319-321); aux = {nb, d0, d1, d2}, no src */
#define SGNN_EXPORT_COPY32 0
#define SGNN_EXPORT_COORDS 1
#define SGNN_EXPORT_CHILDREN 2
#define SGNN_EXPORT_DENSE_CELLS 3
#define SGNN_EXPORT_MAX_SEGS 16
typedef struct SgnnExportSeg {
  const void* src;
  void* dst;
  int64_t n;
  int32_t kind, to_i64;
  int32_t aux[4];
} SgnnExportSeg;
int sgnn_export(const SgnnExportSeg* segs, int32_t n_segs, void* stream);

/* Candidate coordinates of model.py:192-207: out dev [8*n_parent,4]. */
int sgnn_children_coords(const int32_t* parent_coords, int64_t n_parent, int32_t* out, void* stream);

/* ---- a10: GenModel.concat_skip (model.py:338-355): dst[i][col0:col0+c] = src[row of coords[i] in g] or 0. */
int sgnn_concat_skip(const SgnnGrid* g, const float* src, int32_t ld_src, int32_t c,
                     const int32_t* coords, int64_t n, float* dst, int32_t ld_dst, int32_t col0,
                     void* stream);

/* int32 [n,4] -> int64 [n,4] (LongTensor coordinates at the Python boundary). */
int sgnn_coords_to_i64(const int32_t* in, int64_t n, int64_t* out, void* stream);

/* Number of CUDA kernels this library has launched in the calling process (monotonic). */
int64_t sgnn_launch_count(void);

/* Watchdog record of sgnn_conv_forward_tc32_ur: a barrier wait that exceeds ~2 s traps the kernel (the call chain then
 * reports SGNN_E_CUDA) after writing where it stood; out128[0] != 0 when a record exists (128 words, see csrc/conv_ur.cu). */
int sgnn_debug_ur_diag(uint64_t* out128);

/* Measures the sustained 3-register FFMA rate of the device (TFLOP/s): roofline denominator of the fp32 kernels. */
int sgnn_debug_ffma_peak(int iters, double* tflops, void* stream);

int sgnn_version(void);
const char* sgnn_error_string(int code);
int sgnn_last_cuda_error(void);

#ifdef __cplusplus
}
#endif
#endif /* SGNN_B200_H */
