/* rig_grid.h -- C ABI of the B200 grid compositor of librig.so.
 *
 * The reference lays every sample's panels out in one overview image:
 *     compose_grid(W, H, states, candidates, sample_dir, num_options, ...)      /root/reference/src/layout.py:49-192
 *       fit_into_cell(src, cell_size) -> cv2.resize(src, (cell, cell), INTER_AREA)   layout.py:17-31
 *       borders (cv2.rectangle), labels and the '?' query cell (cv2.putText)          layout.py:33-47,131-136,169-174
 * (call site: src/generator.py:517-538).  The only per-sample arithmetic in it is the INTER_AREA down-scale of each 512 x 512
 * panel to its cell; borders, labels and the query cell depend on the layout alone.  The host therefore draws those once per
 * layout with the reference's own calls (an "overlay" image + a per-pixel cell map), and k_grid composes every sample on the
 * device: cell pixels are OpenCV's area resize -- the same float32 multiply / add sequence, in the same order, rounded half to
 * even -- of panels that are already in HBM; everything else is copied from the overlay.
 * Plain C types only; every function returns 0 or a negative RIG_E_* code (include/rig.h).
 */
#ifndef RIG_GRID_H
#define RIG_GRID_H

#include <stddef.h>
#include <stdint.h>

#include "rig.h"

#ifdef __cplusplus
extern "C" {
#endif

/* All pointers are BORROWED host arrays, copied by rig_grid_create. */
typedef struct rig_grid_layout {
    int32_t panel_w, panel_h;     /* source panels (square: panel_w == panel_h) */
    int32_t grid_w, grid_h;       /* overview canvas */
    int32_t cell_size;            /* side of a cell in pixels, < panel size */
    int32_t n_cells;              /* cells that show a panel (<= 64) */
    const int32_t *cell_x;        /* [n_cells] left column of each cell */
    const int32_t *cell_y;        /* [n_cells] top row of each cell */
    const uint8_t *overlay;       /* [grid_h * grid_w * 3] pixels of everything that is not a panel (output channel order) */
    const int8_t *cell_of_pixel;  /* [grid_h * grid_w] cell whose resized panel shows at this pixel, or -1: overlay pixel */
} rig_grid_layout;

typedef struct rig_grid_ctx rig_grid_ctx;

int rig_grid_create(int device, const rig_grid_layout *layout, int max_samples, rig_grid_ctx **out);
void rig_grid_destroy(rig_grid_ctx *ctx);
const char *rig_grid_last_error(const rig_grid_ctx *ctx);

/* Composes n_samples overview images of grid_h * grid_w * 3 bytes.
 *   panels_dev    panel images of panel_h * panel_w * 3 bytes each, on the device (as written by rig_render_*)
 *   panel_index   host array [n_samples * n_cells]: index into panels_dev of the panel shown in each cell, -1: blank (white) cell
 *   out           device (out_is_device = 1, asynchronous on `stream`) or host memory (returns after the copy)          */
int rig_grid_compose(rig_grid_ctx *ctx, const uint8_t *panels_dev, const int32_t *panel_index, int n_samples, uint8_t *out,
                     int out_is_device, void *stream);
int rig_grid_sync(rig_grid_ctx *ctx);

typedef struct rig_grid_stats {
    uint64_t kernel_launches;
    double compose_ms;           /* summed CUDA-event time of the compose kernels since rig_grid_set_profiling(ctx, 1) */
    uint64_t timed_samples;
} rig_grid_stats;
int rig_grid_get_stats(rig_grid_ctx *ctx, rig_grid_stats *out);
int rig_grid_set_profiling(rig_grid_ctx *ctx, int on);

#ifdef __cplusplus
}
#endif
#endif /* RIG_GRID_H */
