/* rig_png.h -- C ABI of the device-side PNG writer of librig.so.
 *
 * The reference turns every rendered image into a PNG file on the host:
 *     save_image(path, img) -> cv2.imwrite(path, img)                  /root/reference/src/utils.py:18-21
 *     (call sites: src/generator.py:342,390,415,501,538)
 *     fig.savefig(save_path, dpi=dpi, ...)  (matplotlib -> libpng)     /root/reference/multigraph_generation/generator.py:256
 * At GPU render rates that is the end-to-end bottleneck twice over: host zlib, and before it the PCIe copy of 786 KB of
 * raw pixels per image.  These entry points encode the images where they are rendered: filter (Sub), LZ77 restricted to
 * zero runs, Huffman coding, Adler-32 and CRC-32 all run in CUDA kernels, and only the finished PNG byte streams (typically
 * 3-12 KB per 512 x 512 image of this data set, about what libpng writes) cross PCIe.  The streams are standard PNG (8-bit RGB
 * or RGBA, one IDAT chunk, a zlib stream of one dynamic-Huffman block whose code the context builds once from a model of this
 * data set's token statistics): any decoder reads them; they are NOT byte-identical to what libpng would write, and the
 * parity criterion is that they decode to exactly the rendered pixels.
 * Plain C types only; every function returns 0 or a negative RIG_E_* code (include/rig.h).
 */
#ifndef RIG_PNG_H
#define RIG_PNG_H

#include <stddef.h>
#include <stdint.h>

#include "rig.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rig_png_ctx rig_png_ctx;

/* Context on CUDA device `device` for W x H x 3 images (W a multiple of 16), at most max_images per call.
 * channels = 3: RGB files (8-bit colour type 2, what cv2.imwrite writes for a 3-channel array);
 * channels = 4: RGBA files (colour type 6) with alpha 255 everywhere -- the format fig.savefig writes, Agg's canvas being RGBA and
 *               opaque (multigraph_generation/generator.py:256).  The input images are 3 bytes per pixel either way.           */
int rig_png_create(int device, int W, int H, int max_images, int channels, rig_png_ctx **out);
void rig_png_destroy(rig_png_ctx *ctx);
const char *rig_png_last_error(const rig_png_ctx *ctx);

/* Upper bound of the bytes n images can encode to (for sizing `out`). */
size_t rig_png_bound(const rig_png_ctx *ctx, int n);

/* Encodes n images of H*W*3 bytes at imgs_dev (device memory, as written by rig_render_* / rig_figure_render_*; `layout`
 * says whether their channel order is RGB or BGR) into n PNG files stored back to back.
 *   offsets_host  [n + 1] uint64, host memory: file i is out[offsets[i] .. offsets[i+1])
 *   out_is_device = 0: `out` is host memory (pinned for full speed) of at least out_capacity bytes; the call returns after
 *                      the copy.  RIG_E_CAPACITY if the files do not fit.
 *   out_is_device = 1: `out` is device memory; the call returns after the kernels have finished (the sizes are needed on
 *                      the host).
 * `stream` (a cudaStream_t or NULL) is the stream the images were rendered on; encoding is ordered after it.           */
int rig_png_encode(rig_png_ctx *ctx, const uint8_t *imgs_dev, int n, int layout, uint8_t *out, size_t out_capacity,
                   uint64_t *offsets_host, int out_is_device, void *stream);

/* The same in two steps, for pipelines: rig_png_encode_begin runs the kernels (it returns once the file sizes are known and the
 * source images have been read) and starts the copy of the files to out_host on a second stream; rig_png_encode_end waits for
 * that copy and hands out the offsets.  Two batches may be in flight (begin, begin, end, end, ...), so the copy of one batch
 * overlaps the kernels of the next.  begin returns the slot (0 or 1) to pass to end, or a negative RIG_E_* code.             */
int rig_png_encode_begin(rig_png_ctx *ctx, const uint8_t *imgs_dev, int n, int layout, uint8_t *out_host, size_t out_capacity,
                         void *stream);
int rig_png_encode_end(rig_png_ctx *ctx, int slot, uint64_t *offsets_host);

/* Makes the encoder's stream wait for `event` (a cudaEvent_t, e.g. from rig_figure_fence) before the next rig_png_encode runs. */
int rig_png_wait_event(rig_png_ctx *ctx, void *event);

/* Optional hint for the NEXT rig_png_encode / rig_png_encode_begin only: flags_dev[n][ceil(H / rows_per_band)] bytes in device
 * memory, 1 = every pixel of those rows is white (255, 255, 255), as rig_figure_blank_bands reports.  Rows so marked are encoded
 * without being read.  A flag of 0 promises nothing; a wrong 1 gives a file in which those rows may be white.  NULL clears the hint.           */
int rig_png_set_blank_hint(rig_png_ctx *ctx, const uint8_t *flags_dev, int rows_per_band);

/* A device buffer of at least n * H * W * 3 bytes owned by the context (grown on demand, 16-byte aligned, valid until the
 * next rig_png_staging call or rig_png_destroy): a place to render into before encoding, for callers that have no device
 * allocator of their own.                                                                                             */
int rig_png_staging(rig_png_ctx *ctx, int n, uint8_t **out_dev);

typedef struct rig_png_stats {
    uint64_t kernel_launches;
    uint64_t images, png_bytes;      /* totals since creation */
    double encode_ms;                /* summed CUDA-event time of the encode kernels since rig_png_set_profiling(ctx, 1) */
    uint64_t timed_images;
} rig_png_stats;
int rig_png_get_stats(rig_png_ctx *ctx, rig_png_stats *out);
int rig_png_set_profiling(rig_png_ctx *ctx, int on);

/* Debug/test hook, needs no GPU: the Huffman code a context for `channels`-channel files uses.  len_out[286]: code lengths of
 * the literal/length symbols; hdr_out[96] / *hdr_bits_out: the deflate block header (BFINAL = 1, BTYPE = 10, code description) as
 * an LSB-first bit string; lut_out[516]: the device token tables (code bits | number of bits << 24): [v] literal v, [256 + l]
 * match of length l = 3..258 at distance 1 (length code, extra bits, the one-bit distance code), [515] end of block.          */
int rig_png_debug_code(int channels, uint8_t *len_out, uint32_t *hdr_out, uint32_t *hdr_bits_out, uint32_t *lut_out);

#ifdef __cplusplus
}
#endif
#endif /* RIG_PNG_H */
