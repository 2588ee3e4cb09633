/*
 * wtb200.h -- C ABI of the B200-native fast wavelet transform (libwtb200.so).
 *
 * This is the drop-in boundary for the ONE hot path of v0lta/PyTorch-Wavelet-Toolbox
 * (ptwt): the multi-level analysis / synthesis filter bank and the boundary-filter
 * matrix FWT.  ptwt has no FFI of its own (it is pure Python that calls torch ops);
 * each entry point below therefore names the reference *Python* body it replaces:
 *
 *   wt_dwt_fwd      <- the level loop of wavedec / wavedec2 / wavedec3
 *                      src/ptwt/conv_transform.py:133-141   (F.pad -> conv1d(stride=2) -> split)
 *                      src/ptwt/conv_transform_2.py:142-149 (F.pad -> conv2d(stride=2) -> split)
 *                      src/ptwt/conv_transform_3.py:122-141 (F.pad -> conv3d(stride=2) -> split)
 *   wt_dwt_inv      <- the level loop of waverec / waverec2 / waverec3
 *                      src/ptwt/conv_transform.py:184-199   (stack -> conv_transpose1d -> crop)
 *                      src/ptwt/conv_transform_2.py:208-249
 *                      src/ptwt/conv_transform_3.py:191-249
 *   wt_matrix_fwd   <- MatrixWavedec.__call__ level loop, src/ptwt/matmul_transform.py:409-425
 *                      (odd-length pad -> torch.sparse.mm(A_level, lo) -> split)
 *   wt_matrix_inv   <- MatrixWaverec.__call__ level loop, src/ptwt/matmul_transform.py:682-699
 *                      (cat -> torch.sparse.mm(S_level, .) -> trim)
 *   wt_matrix_axis_fwd / wt_matrix_axis_inv
 *                   <- the per-axis products of the separable MatrixWavedec2/3 and MatrixWaverec2/3,
 *                      src/ptwt/matmul_transform_2.py:514-531, :797-806; matmul_transform_3.py:255-262, :476-479
 *
 * Conventions
 *   - plain C, no C++ / torch types; every function returns 0 on success, a negative
 *     WT_E* code for a bad argument, or a positive cudaError_t.  wt_last_error()
 *     returns a thread-local message for the last failure on this thread.
 *   - all device buffers are owned by the caller (PyTorch's caching allocator on the
 *     Python side).  The library never allocates or frees device memory in the
 *     device-pointer entry points, never synchronises the device, and keeps no
 *     reference to any argument after it returns.  Work is enqueued on `stream`
 *     (a cudaStream_t passed as void*).
 *   - process-wide state the library DOES keep: (1) one auxiliary non-blocking stream and two events per
 *     device, created on first use and never destroyed: large 2-D analyses (batch >= 16, >= 2^27 samples,
 *     >= 2 levels) run the second half of the batch on it, forked from / joined back into `stream` with
 *     those events, so the call still behaves as if everything ran in `stream` order (no host
 *     synchronisation; disabled while `stream` is being captured into a CUDA graph, or with the switch
 *     NO_AUX_STREAM); (2) the per-kernel dynamic shared-memory opt-in (set once per device and kernel);
 *     (3) the launch counter and the tuning switches at the end of this header.
 *   - filter taps are HOST arrays of double in PyWavelets order (un-flipped dec_lo /
 *     dec_hi / rec_lo / rec_hi, reference src/ptwt/_util.py:95-126); they are rounded
 *     to the compute dtype inside and travel as kernel parameters, so concurrent
 *     streams may use different wavelets.
 *   - extents are ordered slow -> fast (dims[ndim-1] is the contiguous axis);
 *     strides are in ELEMENTS.
 *   - sub-band index k in [0, 2^ndim): k = sum_a hi(a) << (ndim-1-a), axis 0 = slowest,
 *     hi(a) = 1 when the high-pass filter was applied along axis a.
 *       3-D: k = 0..7 = lll, llh, lhl, lhh, hll, hlh, hhl, hhh with the first letter on the
 *            slowest axis -- the reference's own order (src/ptwt/_util.py:926-934), i.e.
 *            keys aad, ada, add, daa, dad, dda, ddd for k = 1..7
 *            (src/ptwt/conv_transform_3.py:131-141).
 *       2-D: k = 1 is lo_H hi_W (reference "hl" = vertical), k = 2 is hi_H lo_W (reference
 *            "lh" = horizontal), k = 3 diagonal (src/ptwt/_util.py:901-905,
 *            src/ptwt/conv_transform_2.py:145-149).
 *       1-D: k = 1 is the detail band.
 */
#ifndef WTB200_H
#define WTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WT_VERSION 100 /* 0.1.0 */

/* dtype */
#define WT_F32 0
#define WT_F64 1

/* boundary modes (ptwt BoundaryMode, src/ptwt/constants.py:85; torch names in
 * src/ptwt/_util.py:36-44) */
#define WT_MODE_ZERO 0      /* "zero"      -> F.pad constant 0            */
#define WT_MODE_CONSTANT 1  /* "constant"  -> F.pad replicate             */
#define WT_MODE_REFLECT 2   /* "reflect"   -> F.pad reflect (no edge repeat) */
#define WT_MODE_PERIODIC 3  /* "periodic"  -> F.pad circular              */
#define WT_MODE_SYMMETRIC 4 /* "symmetric" -> _pad_symmetric (edge repeated) */

/* error codes (negative) */
#define WT_EINVAL (-1)   /* malformed argument                                  */
#define WT_ESHAPE (-2)   /* extents inconsistent with the reference's formulae  */
#define WT_EWORKSPACE (-3) /* workspace too small                               */
#define WT_EUNSUPPORTED (-4)

#define WT_MAX_NDIM 3
#define WT_MAX_FILT_LEN 128

/* One decomposition level of the padded transform, as laid out by the caller.
 * dims[] are the coefficient extents of THIS level; they must equal
 * floor((n_prev + L - 1) / 2) per axis (reference _get_pad, src/ptwt/_util.py:198-228).
 * detail bands k = 1 .. 2^ndim-1 live at  details + (k-1)*band_stride ;
 * the approximation band (k = 0) lives at `approx` (the returned cA for the coarsest
 * level, caller-provided scratch for the others).
 * SCRATCH SEMANTICS (analysis): the `approx` buffers of all but the coarsest level must hold
 * `batch` items, but their contents are UNSPECIFIED on return -- the library processes the
 * batch in chunks and reuses the first few item slots so that the intermediate approximations
 * stay resident in L2 and never travel to HBM.  A caller that wants cA_l of an intermediate
 * level runs a transform with `levels = l`. */
typedef struct wt_level {
    void* details;
    void* approx;
    int64_t dims[WT_MAX_NDIM];
    int64_t strides[WT_MAX_NDIM];        /* element strides inside one detail band */
    int64_t approx_strides[WT_MAX_NDIM]; /* element strides inside the approximation band */
    int64_t details_batch_stride;
    int64_t band_stride;
    int64_t approx_batch_stride;
} wt_level;

int wt_version(void);
const char* wt_last_error(void);

/* Number of SMs / whether a usable sm_100 device is current.  Returns 0 and fills
 * the outputs, or a cudaError_t. */
int wt_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Coefficient extent of one level along one axis: floor((n + L - 1) / 2) for even L
 * (general L: (n + padl + padr - L)/2 + 1 with the reference's pad amounts). */
int64_t wt_coeff_len(int64_t n, int filt_len);

/* Bytes of scratch wt_dwt_fwd / wt_dwt_inv need for the given problem (0 is possible: the fused kernels need
 * none).  inverse: 0 analysis, 1 synthesis; bit 1 set (2, 3) asks for the requirement of the GENERAL path whatever a
 * fused kernel covers -- a fused kernel can still decline at launch time (layouts it does not handle); the transform
 * then returns WT_EWORKSPACE and the caller retries with this amount. */
size_t wt_dwt_workspace_bytes(int ndim, int dtype, int levels, int filt_len, int64_t batch,
                              const int64_t* dims, int inverse);

/* Multi-level analysis.  x is [batch, dims...] with element strides x_strides[ndim] and
 * batch stride x_batch_stride.  levels_desc[0] is the FINEST level (level 1),
 * levels_desc[levels-1] the coarsest.  mode: WT_MODE_*. */
int wt_dwt_fwd(int ndim, int dtype, int mode, int levels, int filt_len,
               const double* dec_lo, const double* dec_hi,
               const void* x, int64_t batch, const int64_t* dims,
               const int64_t* x_strides, int64_t x_batch_stride,
               const wt_level* levels_desc,
               void* workspace, size_t workspace_bytes, void* stream);

/* Multi-level synthesis: consumes levels_desc[levels-1] (coarsest; its `approx` is the
 * input cA) down to levels_desc[0]; intermediate reconstructions are written to the
 * `approx` scratch of the next finer level; the final signal goes to y.
 * out_dims[] are the extents of y; per axis they must be 2*c - L + 2 of level 1, and
 * for every level the reconstructed extent may exceed the next finer level's extent
 * by exactly one sample (the reference trims it, src/ptwt/_util.py:231-244). */
int wt_dwt_inv(int ndim, int dtype, int levels, int filt_len,
               const double* rec_lo, const double* rec_hi,
               void* y, int64_t batch, const int64_t* out_dims,
               const int64_t* y_strides, int64_t y_batch_stride,
               const wt_level* levels_desc,
               void* workspace, size_t workspace_bytes, void* stream);

/* Boundary-filter matrix FWT (MatrixWavedec).  Per level l (0 = finest):
 *   n[l]        even length the level operator acts on (after the optional one-sample
 *               pad of an odd input, reference matmul_transform.py:334-341),
 *   padded[l]   1 if the level input had n[l]-1 samples and is extended by one sample
 *               using odd_mode (WT_MODE_*),
 *   nb_top[l], nb_bot[l]  number of orthogonalised boundary rows at the top / bottom of
 *               EACH half (lo and hi); nb = nb_top + nb_bot rows per half, top rows first,
 *   w_left[l], w_right[l] column support of those rows: the first w_left and the last
 *               w_right columns (the QR leaves round-off sized entries of a top row in the
 *               right corner and vice versa; they are kept, like the reference keeps them),
 *   blocks      device array, compute dtype; for each level in order: lo_left
 *               [nb x w_left], lo_right [nb x w_right], hi_left, hi_right (row-major).
 * x is [batch, n0] contiguous rows (n0 = n[0] - padded[0]).  hi_out[l] receives the detail
 * of level l ([batch, n[l]/2], row stride hi_stride[l]); lo_out the coarsest approximation.
 * scratch must hold 2 * batch * (n[0]/2) elements.
 * allow_fused: non-zero lets consecutive unpadded levels run in one kernel that keeps the approximation
 * in shared memory; that kernel cannot reach the cross-corner entries of the boundary rows (a top row's
 * entries in the right window and vice versa), so the caller sets it only when those entries are
 * negligible (the Python layer: <= 1e-13, true for float64 operators). */
int wt_matrix_fwd(int dtype, int levels, int filt_len,
                  const double* dec_lo, const double* dec_hi,
                  const int64_t* n, const int32_t* padded, int odd_mode,
                  const int32_t* nb_top, const int32_t* nb_bot,
                  const int32_t* w_left, const int32_t* w_right, const void* blocks,
                  const void* x, int64_t batch, int64_t x_stride,
                  void* const* hi_out, const int64_t* hi_stride,
                  void* lo_out, int64_t lo_stride,
                  void* scratch, size_t scratch_bytes, int allow_fused, void* stream);

/* MatrixWaverec: the mirror image.  Taps are the FLIPPED reconstruction filters'
 * source, i.e. pass rec_lo / rec_hi un-flipped (reference flips them itself,
 * matmul_transform.py:110-112); blocks are the boundary rows of S^T per level with the
 * same layout as above.  next_len[l] is the number of samples kept from the level-l
 * reconstruction (n[l] or n[l]-1, reference matmul_transform.py:691-699).  allow_fused as in
 * wt_matrix_fwd: non-zero lets groups of untrimmed levels run as one kernel that keeps the
 * intermediate approximations on chip and drops the cross-corner round-off entries. */
int wt_matrix_inv(int dtype, int levels, int filt_len,
                  const double* rec_lo, const double* rec_hi,
                  const int64_t* n, const int64_t* next_len,
                  const int32_t* nb_top, const int32_t* nb_bot,
                  const int32_t* w_left, const int32_t* w_right, const void* blocks,
                  const void* lo_in, int64_t lo_stride,
                  const void* const* hi_in, const int64_t* hi_stride,
                  int64_t batch, void* y, int64_t y_stride,
                  void* scratch, size_t scratch_bytes, int allow_fused, void* stream);

/* One level of the 1-D boundary-wavelet operator along an arbitrary axis of a [outer, n, inner]
 * tensor whose inner index is contiguous -- the building block of the SEPARABLE 2-D / 3-D matrix
 * transforms (reference src/ptwt/matmul_transform_2.py:514-531 "batch_mm(fwt_col_matrix, ...)",
 * src/ptwt/matmul_transform_3.py:255-262 "_batch_dim_mm(mat, lll, dim)").
 *   analysis:  x [outer, n - padded, inner] -> y [outer, n, inner], low-pass rows 0..n/2-1 then
 *              high-pass rows n/2..n-1 along the axis (the reference's "A x, then split");
 *              padded = 1 appends one sample along the axis per odd_mode first.
 *   synthesis: x [outer, n, inner] (lo | hi along the axis) -> y [outer, keep, inner], keep <= n.
 * blocks: device array, the level's boundary rows as in wt_matrix_fwd / wt_matrix_inv
 * (lo_left, lo_right, hi_left, hi_right).  Strides are in elements. */
int wt_matrix_axis_fwd(int dtype, int filt_len, const double* dec_lo, const double* dec_hi,
                       int64_t n, int padded, int odd_mode,
                       int nb_top, int nb_bot, int w_left, int w_right, const void* blocks,
                       const void* x, int64_t outer, int64_t inner,
                       int64_t x_outer_stride, int64_t x_axis_stride,
                       void* y, int64_t y_outer_stride, int64_t y_axis_stride, void* stream);

int wt_matrix_axis_inv(int dtype, int filt_len, const double* rec_lo, const double* rec_hi,
                       int64_t n, int64_t keep,
                       int nb_top, int nb_bot, int w_left, int w_right, const void* blocks,
                       const void* x, int64_t outer, int64_t inner,
                       int64_t x_outer_stride, int64_t x_axis_stride,
                       void* y, int64_t y_outer_stride, int64_t y_axis_stride, void* stream);

/* Gradient of one transform level with respect to the filter taps along the contiguous axis (learnable wavelets:
 * the reference's filters are nn.Parameters, src/ptwt/wavelets_learnable.py:167-189, and enter its convolutions through
 * src/ptwt/_util.py:129-141).  out[k * L + t] = sum_{r, i} coeff_k[r, i] * sig[r, 2 i + t + 2 - L]  (k = 0 lo, 1 hi;
 * samples outside [0, n) count as zero), accumulated in float64 into the DEVICE array out[2 * L] (zeroed here).
 *   analysis level (zero extension of the explicitly extended input x):  coeff = upstream gradient of the band,
 *       sig = x;   d dec_k[m] = out[k][L - 1 - m]
 *   synthesis level: coeff = the band, sig = upstream gradient of the cropped output;  d rec_k[t] = out[k][t]
 * rows of coeff_* are coeff_stride elements apart (m coefficients each), rows of sig sig_stride apart (n samples). */
int wt_tap_corr(int dtype, int filt_len, const void* coeff_lo, const void* coeff_hi, int64_t coeff_stride,
                const void* sig, int64_t sig_stride, int64_t rows, int64_t m, int64_t n, double* out, void* stream);

/* Counters for bench.py's gpu_launches claim: kernels launched by this library on this
 * process since the last reset. */
uint64_t wt_launch_count(void);
void wt_launch_count_reset(void);

/* Tuning / test switches (NOT part of the drop-in surface).  Every switch is also read ONCE from the
 * environment variable WTB200_<NAME> when the library is first used; no transform call reads the environment.
 * Names: DISABLE_FUSED (general kernels only), NO_WPAIR (one launch per 2-D analysis level), WPAIR_SEG,
 * WPAIR_MIN, WPAIR_DEEP, CHUNK, STREAMS, NO_AUX_STREAM, ENABLE_PAIR, MEGA, ... (csrc/knobs.cuh).
 * wt_set_knob returns 0 or WT_EINVAL for an unknown name; wt_get_knob returns 1 and the value when the switch
 * is set, 0 when it is not, WT_EINVAL for an unknown name. */
int wt_set_knob(const char* name, long long value);
int wt_unset_knob(const char* name);
int wt_get_knob(const char* name, long long* value);

#ifdef __cplusplus
}
#endif
#endif /* WTB200_H */
