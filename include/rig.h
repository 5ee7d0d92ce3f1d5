/* rig.h -- C ABI of the B200 panel rasteriser + near-duplicate filter (librig.so).
 *
 * The reference (MrGu11/reasoning_image_generation) has no FFI; its render stage is the Python method
 *     RPMGenerator._render_elements_to_canvas(W, H, bg_color, elements, use_grid, grid_size)
 *         -> uint8[H, W, 3] BGR                       /root/reference/src/generator.py:66-124
 * which loops Shape.draw() over the element dicts     /root/reference/src/shapes.py:209-667.
 * These entry points are what a ctypes binding of that method calls instead.
 * Plain C types only; every function returns 0 or a negative RIG_E_* code; no exception, abort or
 * CPU fallback ever happens behind this boundary -- if no CUDA device is usable the call fails.
 *
 * Threading: a context may be used from any thread, one call at a time (the reference calls the
 * renderer from a watchdog daemon thread, generator.py:192-197).  The CUDA context is created by
 * rig_create(), never at library load, so the library can be loaded before a fork().
 */
#ifndef RIG_H
#define RIG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIG_OK 0
#define RIG_E_INVALID (-1)      /* bad argument */
#define RIG_E_CUDA (-2)         /* CUDA runtime error; see rig_last_error */
#define RIG_E_CAPACITY (-3)     /* batch larger than the context was created for, or (rig_sync) record arena overflow */
#define RIG_E_UNSUPPORTED (-4)  /* valid for the reference but outside this build's limits */
#define RIG_E_NODEVICE (-5)     /* no usable CUDA device */

#define RIG_LAYOUT_RGB 0
#define RIG_LAYOUT_BGR 1        /* what the reference returns (cv2 order) */

/* shape kinds, in the order of Shape.__init__'s `supported` tuple (shapes.py:201-202) */
enum rig_kind {
    RIG_SQUARE = 0, RIG_CIRCLE, RIG_TRIANGLE, RIG_DIAMOND, RIG_STAR, RIG_PENTAGON, RIG_HEXAGON,
    RIG_PLUS, RIG_HEART, RIG_CRESCENT, RIG_ROUNDED_SQUARE, RIG_KIND_COUNT
};

/* Struct-of-arrays scene description of a batch of panels.  All pointers are BORROWED host arrays
 * that must stay valid until the call returns.  Element i of panel p has
 * panel_elem_begin[p] <= i < panel_elem_begin[p+1]; elements are drawn in index order.
 * Field meaning follows the element dict the reference renderer reads (generator.py:87-94):      */
typedef struct rig_scene {
    int32_t n_panels;
    int32_t n_elements;
    int32_t grid_size;                 /* generator.py:73 (default 3) */
    uint8_t bg_rgb[3];                 /* cfg.bg_color, one per batch (config.py:29) */
    uint8_t reserved;
    const uint32_t *panel_elem_begin;  /* [n_panels + 1] */
    const uint8_t *panel_flags;        /* [n_panels] bit0 = use_grid (snap centres, draw grid lines) */
    const uint8_t *kind;               /* [n_elements] enum rig_kind                el['kind']        */
    const uint8_t *fill;               /* [n_elements] 0/1                          el['fill']        */
    const uint8_t *flip;               /* [n_elements] bit0 horizontal, bit1 vertical; may be NULL
                                          (el['flip_mode'], which the reference never sets: SURVEY B5) */
    const int32_t *stroke;             /* [n_elements]                              int(el['stroke_width']) */
    const int32_t *size;               /* [n_elements]                              int(el['size'])   */
    const int32_t *angle;              /* [n_elements] degrees, clockwise positive  int(el['angle'])  */
    const double *cx;                  /* [n_elements] raw centre as float          el['center'][0]   */
    const double *cy;                  /* [n_elements]                              el['center'][1]   */
    const uint8_t *rgb;                /* [3 * n_elements] fill colour R,G,B        el['color'] (after the
                                          rand_color() draw for colourless elements)        */
} rig_scene;

typedef struct rig_ctx rig_ctx;

/* Creates a context on CUDA device `device` for W x H panels, at most `max_panels` panels and
 * `max_elements` elements per call.  W must be a multiple of 16 (16-byte aligned rows).             */
int rig_create(int device, int W, int H, int max_panels, int max_elements, rig_ctx **out);
void rig_destroy(rig_ctx *ctx);
/* Message of the last failing call on this context (ctx == NULL: of the last failing rig_create). */
const char *rig_last_error(const rig_ctx *ctx);

/* Replaces the body of _render_elements_to_canvas for a whole batch: uploads the scene, rasterises,
 * and writes n_panels images of H*W*3 bytes, panel-major, to `out`.
 *   out_is_device = 0: `out` is host memory (pageable or pinned); the call returns after the copy.
 *   out_is_device = 1: `out` is device memory on the context's device; the call returns after the
 *                      kernels are enqueued on `stream` (a cudaStream_t, or NULL for the context's
 *                      own stream); use rig_sync() before reading.
 *   out_layout: RIG_LAYOUT_RGB or RIG_LAYOUT_BGR.                                                  */
int rig_render_batch(rig_ctx *ctx, const rig_scene *scene, uint8_t *out, int out_is_device, int out_layout,
                     void *stream);

/* Two-step form used when the scene stays resident in HBM (bench.py `value`):
 * rig_scene_upload packs + copies the scene to the device once; rig_render_resident rasterises the
 * resident scene into device memory.                                                               */
int rig_scene_upload(rig_ctx *ctx, const rig_scene *scene, int out_layout, void *stream);
int rig_render_resident(rig_ctx *ctx, uint8_t *out_device, void *stream);

/* The same, with the perceptual hash of every panel (see rig_hash_batch) accumulated inside the rasteriser while each
 * band of rows is still in shared memory: hashes_device receives n_panels * 4 uint64, bit-identical to
 * rig_hash_batch(out_device), without re-reading the pixels from HBM.                                                  */
int rig_render_resident_hashed(rig_ctx *ctx, uint8_t *out_device, uint64_t *hashes_device, void *stream);

/* Waits for the context's streams.  Also the place where an asynchronous (device-destination / resident) render reports a
 * draw-record arena overflow: RIG_E_CAPACITY, the pixels of that render are invalid, and the context's chunk budget has been
 * halved for good (it is not reset by the next upload), so rendering the same scene again succeeds.  Chunks are sized from an
 * upper bound that includes every wrap-around copy an element can have, so this is a defensive path, not an expected one.    */
int rig_sync(rig_ctx *ctx);

/* Counters of the last render on this context (for bench.py's gpu_launches / diagnostics). */
typedef struct rig_stats {
    uint64_t kernel_launches;   /* kernels launched by this library since rig_create */
    uint32_t records;           /* draw records produced by the last render */
    uint32_t aux_units;         /* 16-byte aux units produced by the last render */
    uint32_t overflow;          /* 1 if the last render ran out of record space (result invalid) */
    uint32_t timed_launches;    /* k_raster launches covered by raster_ms / expand_ms */
    double expand_ms;           /* summed CUDA-event time of k_expand since rig_set_profiling(ctx, 1) */
    double raster_ms;           /* summed CUDA-event time of k_raster since rig_set_profiling(ctx, 1) */
    uint64_t timed_panels;      /* panels rendered by those launches */
} rig_stats;
int rig_get_stats(rig_ctx *ctx, rig_stats *out);
/* on != 0: bracket every k_expand / k_raster launch with CUDA events on the launching stream and sum
 * their durations (resolved at rig_sync); turning it on resets the sums.                            */
int rig_set_profiling(rig_ctx *ctx, int on);

/* ---- near-duplicate filter ---------------------------------------------------------------------
 * The reference has no cross-sample duplicate filter (its README mentions one; the code has none --
 * SURVEY section 0, D2/D3), so there is no reference interface to cite: the semantics are defined by
 * this repository (oracle/dedup_oracle.py) and are "parity unpinned".
 *   hash : Y = (77R + 150G + 29B) >> 8, block sums over a 16 x 16 grid, bit b = (256 * S_b > sum S)
 *          -> rig_hash_words() = 4 uint64 per image, block b in word b/64, bit b%64
 *   drop : sample j is dropped iff some i < j has popcount(h_i xor h_j) <= threshold               */
int rig_hash_words(void);
/* imgs_dev: n images of H*W*3 bytes on the device (as written by rig_render_*), hashes_dev: n*4 uint64 */
int rig_hash_batch(rig_ctx *ctx, const uint8_t *imgs_dev, int n, int layout, uint64_t *hashes_dev, void *stream);
/* Scans rows [row_begin, row_end) of the upper-triangular pair matrix over all n_total hashes and
 * sets drop_flags_dev[j] = 1 for every hit (i, j), i < j.  Flags are only ever set, never cleared,
 * so ranks that own disjoint row blocks can share or max-reduce one flag array.  hits_dev (may be
 * NULL) accumulates the number of hit pairs found.                                                 */
int rig_dedup(rig_ctx *ctx, const uint64_t *all_hashes_dev, uint64_t n_total, uint64_t row_begin, uint64_t row_end,
              int threshold, uint8_t *drop_flags_dev, uint64_t *hits_dev, void *stream);

/* Measures the throughput of the integer pipe the scan is bound by, on this device: which = 0 POPC, 1 LOP3 (ALU pipe);
 * lane-operations per second over all SMs (k_int_pipe_peak, dependent chains x 64 warps per SM, best of 3).  The denominator
 * of the scan's roofline (SURVEY 8d asks for a measured, not an assumed, POPC peak).                                        */
int rig_int_pipe_peak(rig_ctx *ctx, int which, double *ops_per_second);

/* Page-locked host memory for scene / output buffers (plain cudaMallocHost / cudaFreeHost). */
int rig_pinned_alloc(size_t bytes, void **out);
void rig_pinned_free(void *p);

/* Debug/test hook: copies the shape tables the device prologue uses (see rig_shapes.h) to `out`
 * (`size` must equal the value returned when out == NULL).  Needs no GPU.                          */
size_t rig_debug_shape_tables(void *out, size_t size);

/* Debugging aid, off unless the environment has RIG_TIMELINE set when the library is loaded: every kernel launch and large copy
 * of every context in the process is bracketed by two CUDA events; rig_timeline_dump synchronises the device and writes
 * "name,stream,start_ms,end_ms" (relative to the first event) to `path`.  tools/e2e_timeline.py reads it.                    */
int rig_timeline_dump(const char *path);

#ifdef __cplusplus
}
#endif
#endif /* RIG_H */
