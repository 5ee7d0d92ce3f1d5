/* rig_figure.h -- C ABI of the B200 figure rasteriser (librig.so), the render stage of the reference's multigraph path.
 *
 * The reference (MrGu11/reasoning_image_generation) has no FFI.  Its multigraph generator builds matplotlib artists and
 * rasterises them with
 *     GeometryGenerator._save_figure(fig, save_path, dpi) -> fig.savefig(...)
 *                                                /root/reference/multigraph_generation/generator.py:250-259
 * i.e. matplotlib's Agg backend, which receives one RendererAgg.draw_path(gc, path, transform, rgbFace) call per artist
 * (patches: multigraph_generation/shapes.py:21-72 styled by style.py:29-66; lines: single_variants.py:285-288,313-333,
 * 352-374; occlusion masks: single_variants.py:552-633).  A batch of such draw_path argument lists is what these entry
 * points take instead (see INTEGRATION.md for the matplotlib-side walker).  Plain C types only; every function returns 0
 * or a negative RIG_E_* code (include/rig.h); no exception, abort or CPU rendering fallback exists behind this boundary.
 *
 * Pipeline: the artist arrays are copied to HBM as they are; k_figure_path runs matplotlib's path converters, Agg's stroker and
 * the rasteriser's clipper on the device (double precision; one thread per artist) and leaves polygon edges in 24.8 fixed
 * point; k_figure accumulates Agg's coverage cells per band of rows in shared memory, blends with matplotlib's
 * non-premultiplied blender and writes each band with one TMA bulk store.  The same geometry code also exists as a host
 * pipeline (host threads, the libm the CPU renderer uses): it is the definition the device is held to, and the library runs it
 * instead whenever the device cannot vouch for bit-identical integers (a cos / sin / atan2 value of a round join or cap
 * that lands within the two libms' error of a rounding boundary) or a path exceeds the device's fixed scratch.
 */
#ifndef RIG_FIGURE_H
#define RIG_FIGURE_H

#include <stddef.h>
#include <stdint.h>

#include "rig.h"

#ifdef __cplusplus
extern "C" {
#endif

/* matplotlib Path codes (matplotlib.path.Path) */
#define RIG_PATH_STOP 0
#define RIG_PATH_MOVETO 1
#define RIG_PATH_LINETO 2
#define RIG_PATH_CURVE3 3
#define RIG_PATH_CURVE4 4
#define RIG_PATH_CLOSEPOLY 79

/* GraphicsContext enums in matplotlib's order (CapStyle, JoinStyle, snap) */
#define RIG_CAP_BUTT 0
#define RIG_CAP_PROJECTING 1
#define RIG_CAP_ROUND 2
#define RIG_JOIN_MITER 0
#define RIG_JOIN_ROUND 1
#define RIG_JOIN_BEVEL 2
#define RIG_SNAP_FALSE 0
#define RIG_SNAP_TRUE 1
#define RIG_SNAP_AUTO 2

/* Struct-of-arrays artist list of a batch of figures.  All pointers are BORROWED host arrays that must stay valid until
 * the call returns.  Artist a of image i has image_artist_begin[i] <= a < image_artist_begin[i+1]; artists are drawn in
 * index order (matplotlib: sorted by zorder, then insertion).  Each artist is one draw_path call:                      */
typedef struct rig_figure_scene {
    int32_t n_images;
    int32_t n_artists;
    int32_t n_vertices;
    int32_t reserved;
    const uint32_t *image_artist_begin;   /* [n_images + 1] */
    const uint32_t *artist_vertex_begin;  /* [n_artists + 1] */
    const double *vertices;               /* [2 * n_vertices] path.vertices, in the artist's own coordinates */
    const uint8_t *codes;                 /* [n_vertices] path.codes (RIG_PATH_*) */
    const double *affine;                 /* [6 * n_artists] transform.to_values(): a, b, c, d, e, f with
                                             x' = a x + c y + e, y' = b x + d y + f -- display pixels, origin bottom-left */
    const double *linewidth;              /* [n_artists] gc line width in PIXELS (points * dpi / 72); 0: no stroke */
    const uint8_t *cap;                   /* [n_artists] RIG_CAP_* */
    const uint8_t *join;                  /* [n_artists] RIG_JOIN_* */
    const uint8_t *snap;                  /* [n_artists] RIG_SNAP_* */
    const uint8_t *edge_rgba;             /* [4 * n_artists] gc colour as rgba8 (agg: uround(c * 255)), non-premultiplied */
    const uint8_t *face_rgba;             /* [4 * n_artists] rgbFace as rgba8 */
    const uint8_t *has_face;              /* [n_artists] 0: rgbFace is None */
    const double *clip;                   /* [4 * n_artists] gc clip rectangle x0, y0, x1, y1 (display), or NULL */
    const uint8_t *has_clip;              /* [n_artists] 0: no clip rectangle (whole canvas); may be NULL */
} rig_figure_scene;

typedef struct rig_figure_ctx rig_figure_ctx;

/* Context on CUDA device `device` for W x H figures (the canvas savefig paints opaque white first), at most max_images
 * images and max_artists artists per call.  W must be a multiple of 16 and W, H <= 3072 (the reference's default
 * savefig, 8 in x 300 dpi, is 2400 x 2400).                              */
int rig_figure_create(int device, int W, int H, int max_images, int max_artists, rig_figure_ctx **out);
void rig_figure_destroy(rig_figure_ctx *ctx);
const char *rig_figure_last_error(const rig_figure_ctx *ctx);

/* Replaces fig.savefig's rasterisation for a whole batch: n_images images of H*W*3 bytes (the RGB of Agg's RGBA buffer;
 * its alpha is 255 everywhere), image-major, to `out`.  out_is_device / out_layout / stream as in rig_render_batch.     */
int rig_figure_render_batch(rig_figure_ctx *ctx, const rig_figure_scene *scene, uint8_t *out, int out_is_device, int out_layout,
                            void *stream);

/* Where the geometry half of draw_path (transform, snap, curve flattening, stroking, clipping) runs.  RIG_GEOM_DEVICE is the
 * default (environment RIG_FIG_GEOM=host overrides it at context creation); the results are identical.                 */
#define RIG_GEOM_HOST 0
#define RIG_GEOM_DEVICE 1
int rig_figure_set_geometry(rig_figure_ctx *ctx, int mode);

/* Two-step form for scenes that stay resident in HBM (bench.py `value`): upload copies the artist arrays to the device and
 * runs the geometry once (polygon edges resident).  render_resident rasterises those edges into device memory (k_figure
 * only); draw_resident is the whole draw_path for every artist of the resident scene -- k_figure_path from the artist arrays,
 * then k_figure -- and needs a scene the device geometry accepted (RIG_E_UNSUPPORTED otherwise).  Both are asynchronous;
 * rig_figure_sync waits and reports a draw_resident whose geometry kernel raised a status.                               */
int rig_figure_upload(rig_figure_ctx *ctx, const rig_figure_scene *scene, int out_layout, void *stream);
int rig_figure_render_resident(rig_figure_ctx *ctx, uint8_t *out_device, void *stream);
int rig_figure_draw_resident(rig_figure_ctx *ctx, uint8_t *out_device, void *stream);
int rig_figure_sync(rig_figure_ctx *ctx);
/* The context's own stream (a cudaStream_t): what asynchronous calls run on when `stream` is NULL.  Hand it to a consumer of the
 * rendered images (e.g. rig_png_encode) to order it after the render without a host synchronisation.                     */
void *rig_figure_stream(rig_figure_ctx *ctx);
/* Records an event (a cudaEvent_t from a ring of 8, so the handle stays valid for the next 7 calls) behind everything enqueued
 * on the context's stream so far, e.g. right after an asynchronous rig_figure_render_batch: a consumer waits on exactly that
 * render (rig_png_wait_event), not on whatever another thread enqueues later.                                             */
int rig_figure_fence(rig_figure_ctx *ctx, void **event_out);

/* Optional side product of renders into device memory (rig_figure_render_batch with out_is_device = 1, rig_figure_render_resident),
 * switched on with rig_figure_set_blank_flags(ctx, 1) (off by default: the byte store per band costs k_figure about 2 %):
 * one byte per band of *rows_per_band_out rows, image-major ([n_images][ceil(H / rows)]), in device memory; 1 = no artist reaches
 * the band, every pixel in it is the white background.  Written by k_figure on the render's stream; rig_figure_blank_bands gives
 * the flags of the last render, valid until the fourth next render (NULL if that render wrote none).  A consumer of the images
 * (rig_png_set_blank_hint) can skip those rows without reading them.                                                      */
int rig_figure_set_blank_flags(rig_figure_ctx *ctx, int on);
int rig_figure_blank_bands(rig_figure_ctx *ctx, const uint8_t **flags_dev_out, int *rows_per_band_out);

typedef struct rig_figure_stats {
    uint64_t kernel_launches;   /* kernels launched by this context since creation */
    uint64_t edges;             /* polygon edges of the resident scene */
    uint64_t layers;            /* rasteriser passes (faces + strokes) of the resident scene */
    uint32_t timed_launches;    /* k_figure launches covered by cover_ms */
    uint32_t reserved;
    double cover_ms;            /* summed CUDA-event time of k_figure since rig_figure_set_profiling(ctx, 1) */
    uint64_t timed_images;
    double prepare_ms;          /* host time of the last upload: packing the artist arrays, or the host path pipeline */
    double geom_ms;             /* summed CUDA-event time of k_figure_path since rig_figure_set_profiling(ctx, 1) */
    uint32_t geom_launches;     /* k_figure_path launches covered by geom_ms */
    uint32_t geom_fallbacks;    /* uploads the device geometry handed back to the host pipeline since creation */
    uint32_t geom_status;       /* status word of the last device geometry run (0: accepted; bits: rig_figure_geom.h GEOM_*) */
    uint32_t geom_resident;     /* 1: the resident scene came from the device geometry */
} rig_figure_stats;
int rig_figure_get_stats(rig_figure_ctx *ctx, rig_figure_stats *out);
int rig_figure_set_profiling(rig_figure_ctx *ctx, int on);

/* Debug/test hook, needs no GPU: runs the host path pipeline and copies what would be uploaded.  layers_out: 8 uint32 per
 * layer (image, edge_begin, n_edges, rgba, xmin, xmax, ymin, ymax); edges_out: 4 int32 per edge (x1, y1, x2, y2, 24.8 fixed
 * point, y down).  Returns the number of edges, or a negative RIG_E_* code (RIG_E_CAPACITY: more than edge_cap edges).  */
int64_t rig_figure_debug_edges(const rig_figure_scene *scene, int W, int H, uint32_t *layers_out, int32_t *edges_out,
                               int64_t edge_cap, int64_t *n_layers_out);

/* Debug/test hook: the layers and edges of the RESIDENT scene (whichever geometry produced them), copied back from the device
 * in the format of rig_figure_debug_edges (empty passes dropped, edges re-packed in layer order).                        */
int64_t rig_figure_debug_resident(rig_figure_ctx *ctx, uint32_t *layers_out, int32_t *edges_out, int64_t edge_cap,
                                  int64_t *n_layers_out);

#ifdef __cplusplus
}
#endif
#endif /* RIG_FIGURE_H */
