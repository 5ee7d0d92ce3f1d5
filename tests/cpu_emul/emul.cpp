// tests/cpu_emul/emul.cpp -- TEST-ONLY serial driver for the product's per-lane code.
//
// Compiles the very same headers the CUDA kernels use (rig_core.h / rig_shapes.h / rig_expand.h /
// rig_raster.h) with LANES = 1 and runs them on the host, band by band, so that the record
// arithmetic and the band clipping can be compared with the oracle without a GPU.  It is not part
// of librig.so and is never used by the product path.
#include <string.h>
#include <vector>

#include "../../reasoning_image_generation_b200/csrc/rig_expand.h"
#include "../../reasoning_image_generation_b200/csrc/rig_host.h"
#include "../../reasoning_image_generation_b200/csrc/rig_raster.h"

using namespace rig;

struct HostOps {
    const Rec *cur = nullptr;
    Rec slot(int k) const { return cur[k]; }
    Rec slot_part(int k, int lo, int hi) const
    {
        Rec r = rec_nop();
        for (int i = lo; i <= hi; i++) r.w[i] = cur[k].w[i];
        return r;
    }
    int bcast(int v, int) const { return v; }
    int reduce_add(int v) const { return v; }
    int reduce_min(int v) const { return v; }
    int reduce_max(int v) const { return v; }
    void sync() const {}
    void append(RasterScratch &S, int32_t v) const
    {
        if (S.count < 64) S.cross[S.count] = v;
        S.count++;
    }
};
unsigned long long g_exec_count[16];
unsigned long long g_window_loads, g_band_nonempty, g_bands;
struct NoSync {
    void operator()() const {}
};

extern "C" int emul_render(const rig_scene *scene, int W, int H, int band_rows, int layout, uint8_t *out,
                           uint32_t *nrec_out, char *err, size_t errlen)
{
    static ShapeTables T;
    static bool init = false;
    if (!init) { fill_shape_tables(T); init = true; }
    std::vector<DElem> elems(scene->n_elements ? scene->n_elements : 1);
    int rc = pack_elements(*scene, W, H, layout == RIG_LAYOUT_BGR, elems.data(), err, errlen);
    if (rc) return rc;
    const int gs = scene->grid_size > 0 ? scene->grid_size : 3;
    uint32_t bg = layout == RIG_LAYOUT_BGR
                      ? (uint32_t)scene->bg_rgb[2] | (scene->bg_rgb[1] << 8) | (scene->bg_rgb[0] << 16)
                      : (uint32_t)scene->bg_rgb[0] | (scene->bg_rgb[1] << 8) | (scene->bg_rgb[2] << 16);
    uint32_t total = 0;
    for (int p = 0; p < scene->n_panels; p++) {
        std::vector<Rec> recs;
        std::vector<Q16> aux;
        std::vector<ElemHdr> hdrs;
        ExpandScratch ES;
        for (uint32_t i = scene->panel_elem_begin[p]; i < scene->panel_elem_begin[p + 1]; i++) {
            ElemPlan plan = expand_plan<1>(elems[i], T, W, H, 0, ES, NoSync());
            uint32_t r0 = (uint32_t)recs.size(), a0 = (uint32_t)aux.size();
            const uint32_t nsum = summary_slots(plan.nrec);
            recs.resize(r0 + nsum + plan.nrec + 8, rec_nop());
            aux.resize(a0 + plan.naux + 1);
            expand_emit<1>(elems[i], plan, T, W, H, 0, ES, recs.data(), r0 + nsum, aux.data(), a0, NoSync());
            write_summaries<1>(recs.data(), r0, r0 + nsum, plan.nrec, 0);
            ElemHdr h;
            h.rec_begin = r0; h.nrec = plan.nrec; h.ymin = (int16_t)plan.ymin; h.ymax = (int16_t)plan.ymax; h.nsum = nsum;
            hdrs.push_back(h);
        }
        total += (uint32_t)recs.size();
        uint8_t *img = out + (size_t)p * W * H * 3;
        const bool use_grid = scene->panel_flags && (scene->panel_flags[p] & 1);
        RasterScratch RS;
        for (int r0 = 0; r0 < H; r0 += band_rows) {
            Band b;
            b.buf = img + (size_t)r0 * W * 3;
            b.r0 = r0;
            b.r1 = r0 + band_rows - 1 < H - 1 ? r0 + band_rows - 1 : H - 1;
            b.W = W; b.H = H;
            for (int y = b.r0; y <= b.r1; y++)
                for (int x = 0; x < W; x++) put3(band_px(b, x, y), bg);
            for (size_t e = 0; e < hdrs.size(); e++) {
                if (hdrs[e].ymin > b.r1 || hdrs[e].ymax < b.r0) continue;
                const uint32_t *sums = (const uint32_t *)&recs[hdrs[e].rec_begin];
                const uint32_t pb = hdrs[e].rec_begin + hdrs[e].nsum;
                for (uint32_t k = 0; k < hdrs[e].nrec; k++) {
                    const uint32_t sv = sums[k / 32];
                    if ((int16_t)(sv & 0xffff) > b.r1 || (int16_t)(sv >> 16) < b.r0) continue;   // window skipped
                    if (k % 32 == 0 || true) {}
                    HostOps ops;
                    ops.cur = &recs[pb + k];
                    { const Rec &rr = recs[pb + k]; if (rec_ymin(rr) <= b.r1 && rec_ymax(rr) >= b.r0) g_exec_count[rec_type(rr) & 15]++; }
                    exec_record<1>(recs[pb + k], aux.data(), T.filter, b, 0, RS, ops);
                }
            }
            if (use_grid) {
                int16_t gx[16], gy[16];
                for (int g = 1; g < gs && g <= 16; g++) {
                    volatile double qx = (double)(g * W) / (double)gs, qy = (double)(g * H) / (double)gs;
                    gx[g - 1] = (int16_t)nearbyint(qx);
                    gy[g - 1] = (int16_t)nearbyint(qy);
                }
                draw_grid<1>(b, gs - 1, gx, gy, 0);
            }
        }
    }
    if (nrec_out) *nrec_out = total;
    return 0;
}

extern "C" void emul_counts(unsigned long long *out) { for (int i = 0; i < 16; i++) { out[i] = g_exec_count[i]; g_exec_count[i] = 0; } }

extern "C" size_t emul_shape_tables(void *out, size_t size)
{
    if (out && size == sizeof(ShapeTables)) {
        ShapeTables T;
        fill_shape_tables(T);
        memcpy(out, &T, sizeof T);
    }
    return sizeof(ShapeTables);
}

// brute-force checks of the closed forms against the sequential loops they replace
extern "C" int emul_selftest_line8(int dx, int dy)
{
    // sequential error walk
    int err = dx - (dy + dy), m = 0;
    for (int k = 0; k <= dx; k++) {
        if (line8_minor(dx, dy, k) != m) return 1;
        int neg = err < 0;
        err += -(dy + dy) + (neg ? dx + dx : 0);
        m += neg;
    }
    for (int mm = 1; mm <= dy; mm++) {
        int k = line8_first_k(dx, dy, mm);
        if (k > dx) continue;
        if (line8_minor(dx, dy, k) < mm) return 2;
        if (k > 0 && line8_minor(dx, dy, k - 1) >= mm) return 3;
    }
    return 0;
}
