// rig_host.h -- host-side helpers shared by the C ABI (rig_abi.cu) and the CPU emulation used in tests.
#pragma once
#include <stddef.h>
#include "../../include/rig.h"
#include "rig_shapes.h"

namespace rig {

// unit shape tables + OpenCV line tables (rig_tables.cpp)
void fill_shape_tables(ShapeTables &T);

// rig_scene (borrowed SoA) -> packed device elements; applies int() truncation / grid snap of the
// centre (generator.py:97-100, shapes.py:231) and evaluates cos/sin of radians(-angle) with libm.
// Returns RIG_OK or a negative code with a message in err.
int pack_elements(const rig_scene &s, int W, int H, int layout_bgr, DElem *out, char *err, size_t errlen);

}  // namespace rig
