// rig_figure_host.h -- host half of the figure path: artists -> rasteriser layers (edges in 24.8 fixed point).
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/rig_figure.h"
#include "rig_figure_core.h"

namespace rig {

struct FigurePrepared {
    std::vector<FLayer> layers;           // in draw order, image by image
    std::vector<uint32_t> image_layer_begin;   // [n_images + 1]
    std::vector<FEdge> edges;
    std::vector<uint32_t> dir;
};

// Runs matplotlib's converter chain (transform, snap, curve flattening), Agg's stroker and the rasteriser's clipper for
// every artist of images [i0, i1) and appends the layers / edges.  Pure host code, thread safe on disjoint outputs.
// Returns 0 or a RIG_E_* code with a message in err.
int figure_prepare(const rig_figure_scene &s, int W, int H, bool bgr, int i0, int i1, FigurePrepared &out, char *err, int errlen);

}  // namespace rig
