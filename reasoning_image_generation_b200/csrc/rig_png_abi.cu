// rig_png_abi.cu -- the extern "C" boundary declared in include/rig_png.h.
#include <cuda_runtime.h>
#include <sched.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/rig_png.h"
#include "rig_png.cuh"
#include "rig_timeline.h"

using namespace rig;

namespace {
thread_local char g_png_create_err[512] = "";

uint32_t host_multmodp(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
    for (int i = 0; i < 32; i++) {
        if (a & (0x80000000u >> i)) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}
}  // namespace

struct rig_png_ctx {
    int device = 0, W = 0, H = 0, max_images = 0, ch = 3;
    uint32_t ppi = 0;
    int wpc = 4, sm_count = 1, ctas_per_sm = 1;
    int asm_warps = 8;                 // warps (= files) per CTA of k_png_assemble
    size_t smem = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_in = nullptr, ev_asm = nullptr;
    cudaEvent_t ev_sizes = nullptr;    // blocking-sync event: the host thread sleeps while the kernels run
    bool blocking = false;
    // two output slots: the copy of one batch's files to the host overlaps the kernels of the next batch
    cudaEvent_t ev0[2] = {nullptr, nullptr}, ev1[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    uint8_t *d_outs[2] = {nullptr, nullptr};
    size_t out_caps[2] = {0, 0};
    unsigned long long *h_offs[2] = {nullptr, nullptr};   // pinned
    int slot_n[2] = {0, 0};
    unsigned long long *h_counters = nullptr;             // pinned copy of d_counters[0..1]
    bool slot_busy[2] = {false, false};
    unsigned next_slot = 0;
    const uint8_t *hint = nullptr;     // blank-band flags for the next encode only
    int hint_rows = 0;
    PngTables *d_tables = nullptr;
    PngPiece *d_pieces = nullptr;
    uint8_t *d_scratch = nullptr, *d_stage = nullptr;
    size_t scratch_bytes = 0, over_begin = 0, stage_cap = 0;
    unsigned long long *d_sizes = nullptr, *d_offsets = nullptr, *d_counters = nullptr;
    rig_png_stats stats = {};
    bool profiling = false;
    char err[512] = "";
};

namespace {
int pfail(rig_png_ctx *c, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c ? c->err : g_png_create_err, 512, fmt, ap);
    va_end(ap);
    return code;
}
#define PCK(ctx, call)                                                                                     \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return pfail(ctx, RIG_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
}  // namespace

extern "C" {

const char *rig_png_last_error(const rig_png_ctx *ctx) { return ctx ? ctx->err : g_png_create_err; }

size_t rig_png_bound(const rig_png_ctx *c, int n)
{
    if (!c || n <= 0) return 0;
    // literal-only worst case: the longest code per filtered byte, plus the block header and the file framing
    const size_t bits = (size_t)c->H * (1 + (size_t)c->ch * c->W) * PNG_MAX_CODE_BITS + PNG_HDR_WORDS * 32 + 15;
    return (size_t)n * ((bits + 7) / 8 + PNG_FILE_OVERHEAD);
}

// cores this process may run on (its affinity mask: what a container's cpuset or taskset leaves), else the machine's
static unsigned usable_cores()
{
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) return (unsigned)CPU_COUNT(&set);
    return std::max(1u, std::thread::hardware_concurrency());
}

// How the host waits for the GPU: spinning in cudaStreamSynchronize keeps a core busy; sleeping on a blocking-sync event costs a
// wake-up.  Measured end to end (figures, RGBA files): 16 cores per GPU: spin 1.04-1.05 M, sleep 0.87-0.97 M images/s; 4 cores per GPU
// (where the spinning thread takes a quarter of what the renderer's host pipeline needs): spin 0.54-0.56 M, sleep 0.69-0.80 M.
// So: sleep when fewer than 8 cores per rank (cores / LOCAL_WORLD_SIZE); RIG_PNG_BLOCKING_SYNC=0/1 overrides.
static bool png_blocking_wait()
{
    if (const char *e = getenv("RIG_PNG_BLOCKING_SYNC")) return atoi(e) != 0;
    unsigned cores = usable_cores();
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) {
        const int ranks = atoi(e);
        if (ranks > 1 && cores > (unsigned)ranks) cores = cores / (unsigned)ranks;
    }
    return cores < 8u;
}

int rig_png_create(int device, int W, int H, int max_images, int channels, rig_png_ctx **out)
{
    if (!out) return pfail(nullptr, RIG_E_INVALID, "out is NULL");
    *out = nullptr;
    if (W <= 0 || H <= 0 || W % 16 || W > 4096 || H > 4096)
        return pfail(nullptr, RIG_E_INVALID, "unsupported image %dx%d (W %% 16 == 0, <= 4096)", W, H);
    if (max_images <= 0) return pfail(nullptr, RIG_E_INVALID, "max_images must be positive");
    if (channels != 3 && channels != 4) return pfail(nullptr, RIG_E_INVALID, "channels must be 3 (RGB files) or 4 (RGBA files)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return pfail(nullptr, RIG_E_NODEVICE, "no usable CUDA device (%s); this library has no CPU path",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return pfail(nullptr, RIG_E_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    rig_png_ctx *c = new rig_png_ctx();
    c->device = device; c->W = W; c->H = H; c->max_images = max_images; c->ch = channels;
    c->ppi = (uint32_t)((H + PNG_ROWS - 1) / PNG_ROWS);
    auto bail = [&](int code) {
        snprintf(g_png_create_err, sizeof g_png_create_err, "%s", c->err);
        rig_png_destroy(c);
        return code;
    };
#define PCKC(call)                                                                                 \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            pfail(c, RIG_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                  \
            return bail(RIG_E_CUDA);                                                               \
        }                                                                                          \
    } while (0)
    PCKC(cudaSetDevice(device));
    cudaDeviceProp prop;
    PCKC(cudaGetDeviceProperties(&prop, device));
    // warps per CTA: the largest CTA of which three still fit an SM (measured, profiles/r1_notes.md: fewer, larger CTAs lose
    // occupancy to the register file; many small ones pay the per-CTA token tables more often), else two, else one
    {
        const size_t per_warp = png_warp_smem(W, channels), sm_bytes = prop.sharedMemPerMultiprocessor;
        int best = 1;
        for (int want = 3; want >= 1 && best == 1; want--)
            for (int w = 8; w >= 1; w--) {
                const size_t cta = (size_t)w * per_warp + 3072 + 1024;      // + the static token tables + the per-CTA reservation
                if (cta <= prop.sharedMemPerBlockOptin && sm_bytes / cta >= (size_t)want) { best = w; break; }
            }
        c->wpc = best;
        if (const char *ev = getenv("RIG_PNG_WPC")) { const int w = atoi(ev); if (w >= 1 && w <= 8) c->wpc = w; }
    }
    c->smem = c->wpc * png_warp_smem(W, channels);
    c->sm_count = prop.multiProcessorCount;
    c->ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(2048 / (c->wpc * 32), prop.sharedMemPerMultiprocessor / (c->smem + 3072 + 1024)));
    if (c->smem > prop.sharedMemPerBlockOptin) {
        pfail(c, RIG_E_UNSUPPORTED, "image width %d needs %zu B of shared memory per piece", W, c->smem);
        return bail(RIG_E_UNSUPPORTED);
    }
    if (channels == 3) PCKC(cudaFuncSetAttribute(k_png_pieces<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem));
    else PCKC(cudaFuncSetAttribute(k_png_pieces<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem));
    c->blocking = png_blocking_wait();
    PCKC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    PCKC(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    PCKC(cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming));
    PCKC(cudaEventCreateWithFlags(&c->ev_asm, cudaEventDisableTiming));
    PCKC(cudaMallocHost(&c->h_counters, 2 * sizeof(unsigned long long)));
    PCKC(cudaEventCreateWithFlags(&c->ev_sizes, cudaEventDisableTiming | cudaEventBlockingSync));
    for (int i = 0; i < 2; i++) {
        PCKC(cudaEventCreate(&c->ev0[i]));
        PCKC(cudaEventCreate(&c->ev1[i]));
        PCKC(cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming | (c->blocking ? cudaEventBlockingSync : 0)));
        PCKC(cudaMallocHost(&c->h_offs[i], sizeof(unsigned long long) * ((size_t)max_images + 1)));
    }
    PngTables T;
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t v = i;
        for (int k = 0; k < 8; k++) v = (v & 1u) ? (v >> 1) ^ CRC_POLY : v >> 1;
        T.crc[i] = v;
    }
    T.x2n[0] = 1u << 30;
    for (int i = 1; i < 32; i++) T.x2n[i] = host_multmodp(T.x2n[i - 1], T.x2n[i - 1]);
    {
        const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
        memcpy(T.header, sig, 8);
        uint8_t *h = T.header + 8;
        h[0] = 0; h[1] = 0; h[2] = 0; h[3] = 13;
        memcpy(h + 4, "IHDR", 4);
        h[8] = (uint8_t)(W >> 24); h[9] = (uint8_t)(W >> 16); h[10] = (uint8_t)(W >> 8); h[11] = (uint8_t)W;
        h[12] = (uint8_t)(H >> 24); h[13] = (uint8_t)(H >> 16); h[14] = (uint8_t)(H >> 8); h[15] = (uint8_t)H;
        h[16] = 8; h[17] = channels == 4 ? 6 : 2; h[18] = 0; h[19] = 0; h[20] = 0;      // 8-bit RGB / RGBA, deflate, adaptive filtering, no interlace
        uint32_t crc = 0xffffffffu;
        for (int i = 4; i < 21; i++) crc = T.crc[(crc ^ h[i]) & 0xff] ^ (crc >> 8);
        crc ^= 0xffffffffu;
        h[21] = (uint8_t)(crc >> 24); h[22] = (uint8_t)(crc >> 16); h[23] = (uint8_t)(crc >> 8); h[24] = (uint8_t)crc;
    }
    {
        PngCode code;
        if (!png_build_code(channels, code)) {
            pfail(c, RIG_E_UNSUPPORTED, "internal: the Huffman code description does not fit the header area");
            return bail(RIG_E_UNSUPPORTED);
        }
        for (uint32_t v = 0; v < 256; v++) T.lit[v] = png_lut_literal(code, v);
        T.match[0] = T.match[1] = T.match[2] = 0;
        for (uint32_t l = 3; l <= 258; l++) T.match[l] = png_lut_match(code, l);
        T.eob = png_lut_literal(code, 256);
        T.hdr_bits = code.hdr_bits;
        memcpy(T.hdr, code.hdr, sizeof T.hdr);
    }
    while (c->asm_warps > 1 && png_assemble_smem(c->ppi, c->asm_warps) + 2048 > prop.sharedMemPerBlockOptin) c->asm_warps /= 2;
    if (png_assemble_smem(c->ppi, c->asm_warps) > 40 * 1024)
        PCKC(cudaFuncSetAttribute(k_png_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)png_assemble_smem(c->ppi, c->asm_warps)));
    PCKC(cudaMalloc(&c->d_tables, sizeof T));
    PCKC(cudaMemcpy(c->d_tables, &T, sizeof T, cudaMemcpyHostToDevice));
    const size_t npieces = (size_t)max_images * c->ppi;
    PCKC(cudaMalloc(&c->d_pieces, sizeof(PngPiece) * npieces));
    c->over_begin = npieces * PNG_SLOT;
    c->scratch_bytes = 2 * c->over_begin + (1u << 20);
    PCKC(cudaMalloc(&c->d_scratch, c->scratch_bytes));
    PCKC(cudaMalloc(&c->d_sizes, sizeof(unsigned long long) * ((size_t)max_images + 1)));
    PCKC(cudaMalloc(&c->d_offsets, sizeof(unsigned long long) * ((size_t)max_images + 1)));
    PCKC(cudaMalloc(&c->d_counters, sizeof(unsigned long long) * 4));
#undef PCKC
    *out = c;
    return RIG_OK;
}

void rig_png_destroy(rig_png_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    cudaFree(c->d_tables); cudaFree(c->d_pieces); cudaFree(c->d_scratch); cudaFree(c->d_sizes); cudaFree(c->d_offsets);
    cudaFree(c->d_counters); cudaFree(c->d_stage);
    for (int i = 0; i < 2; i++) {
        cudaFree(c->d_outs[i]);
        cudaFreeHost(c->h_offs[i]);
        if (c->ev0[i]) cudaEventDestroy(c->ev0[i]);
        if (c->ev1[i]) cudaEventDestroy(c->ev1[i]);
        if (c->ev_copy[i]) cudaEventDestroy(c->ev_copy[i]);
    }
    if (c->ev_in) cudaEventDestroy(c->ev_in);
    if (c->ev_asm) cudaEventDestroy(c->ev_asm);
    if (c->ev_sizes) cudaEventDestroy(c->ev_sizes);
    cudaFreeHost(c->h_counters);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

// kernels of one batch up to the file offsets; returns the total number of bytes in *total (the host waits for the sizes)
static int png_run_pieces(rig_png_ctx *c, const uint8_t *imgs_dev, int n, int layout, void *stream, int slot, unsigned long long *total,
                          const uint8_t *hint, int hint_rows)
{
    cudaStream_t st = c->stream;
    // order the encoder after the stream the images were rendered on
    PCK(c, cudaEventRecord(c->ev_in, (cudaStream_t)stream));
    PCK(c, cudaStreamWaitEvent(st, c->ev_in, 0));
    PCK(c, cudaMemsetAsync(c->d_counters, 0, 4 * sizeof(unsigned long long), st));
    if (c->profiling) PCK(c, cudaEventRecord(c->ev0[slot], st));
    PngParams P;
    P.imgs = imgs_dev; P.n = n; P.W = c->W; P.H = c->H; P.bgr = layout == RIG_LAYOUT_BGR;
    P.pieces_per_image = c->ppi; P.pieces = c->d_pieces; P.scratch = c->d_scratch;
    P.over_begin = c->over_begin; P.over_cap = c->scratch_bytes - c->over_begin; P.counters = c->d_counters;
    P.tables = c->d_tables;
    P.blank = hint; P.blank_rows = hint_rows; P.blank_bands = hint ? (uint32_t)((c->H + hint_rows - 1) / hint_rows) : 0u;
    const unsigned long long npieces = (unsigned long long)n * c->ppi;
    {
        // persistent warps: every warp prepares the bits of a white row once, then walks pieces with a grid stride
        const unsigned long long want = (npieces + c->wpc - 1) / c->wpc;
        const unsigned grid = (unsigned)std::min<unsigned long long>(want, (unsigned long long)c->sm_count * c->ctas_per_sm);
        TimelineScope tl("k_png_pieces", st);
        if (c->ch == 3) k_png_pieces<3><<<grid, c->wpc * 32, c->smem, st>>>(P);
        else k_png_pieces<4><<<grid, c->wpc * 32, c->smem, st>>>(P);
    }
    k_png_sizes<<<(n + 7) / 8, 256, 0, st>>>(c->d_pieces, n, c->ppi, c->d_tables, c->d_sizes);
    unsigned long long *h_counters = c->h_counters;      // page-locked: k_png_scan writes the offsets and the counters there itself
    k_png_scan<<<1, 1024, 0, st>>>(c->d_sizes, n, c->d_offsets, c->h_offs[slot], c->d_counters, h_counters);
    c->stats.kernel_launches += 3;
    PCK(c, cudaGetLastError());
    // the host needs the sizes before it can go on (how it waits: see png_blocking_wait)
    if (!c->blocking) PCK(c, cudaStreamSynchronize(st));
    else {
        PCK(c, cudaEventRecord(c->ev_sizes, st));
        PCK(c, cudaEventSynchronize(c->ev_sizes));
    }
    if (h_counters[1]) return pfail(c, RIG_E_CAPACITY, "the images are too incompressible for the piece scratch area (%d images); encode fewer per call", n);
    *total = c->h_offs[slot][n];
    return RIG_OK;
}

static int png_run_assemble(rig_png_ctx *c, int n, uint8_t *dev_out, int slot)
{
    PngAssembleParams A;
    A.pieces = c->d_pieces; A.scratch = c->d_scratch; A.offsets = c->d_offsets; A.tables = c->d_tables; A.out = dev_out;
    A.n = n; A.W = c->W; A.H = c->H; A.ch = c->ch; A.pieces_per_image = c->ppi;
    {
        TimelineScope tl("k_png_assemble", c->stream);
        k_png_assemble<<<(n + c->asm_warps - 1) / c->asm_warps, c->asm_warps * 32, png_assemble_smem(c->ppi, c->asm_warps), c->stream>>>(A);
    }
    c->stats.kernel_launches++;
    PCK(c, cudaGetLastError());
    if (c->profiling) PCK(c, cudaEventRecord(c->ev1[slot], c->stream));
    return RIG_OK;
}

static int png_check_args(rig_png_ctx *c, const uint8_t *imgs_dev, int n, int layout, const void *out)
{
    if (n < 0 || (n && (!imgs_dev || !out))) return pfail(c, RIG_E_INVALID, "bad arguments to rig_png_encode");
    if (n > c->max_images) return pfail(c, RIG_E_CAPACITY, "%d images > max_images %d", n, c->max_images);
    if (layout != RIG_LAYOUT_RGB && layout != RIG_LAYOUT_BGR) return pfail(c, RIG_E_INVALID, "bad layout %d", layout);
    if (((uintptr_t)imgs_dev) % 16) return pfail(c, RIG_E_INVALID, "imgs_dev must be 16-byte aligned");
    return RIG_OK;
}

int rig_png_encode_begin(rig_png_ctx *c, const uint8_t *imgs_dev, int n, int layout, uint8_t *out_host, size_t out_capacity, void *stream)
{
    if (!c) return RIG_E_INVALID;
    const uint8_t *hint = c->hint;          // for this call only, whatever its outcome
    const int hint_rows = c->hint_rows;
    c->hint = nullptr;
    int rc = png_check_args(c, imgs_dev, n, layout, out_host);
    if (rc) return rc;
    PCK(c, cudaSetDevice(c->device));
    const int slot = (int)(c->next_slot & 1u);
    if (c->slot_busy[slot]) return pfail(c, RIG_E_INVALID, "both output slots are in flight: call rig_png_encode_end first");
    c->slot_n[slot] = n;
    if (n == 0) { c->h_offs[slot][0] = 0; c->slot_busy[slot] = true; c->next_slot++; return slot; }
    unsigned long long total = 0;
    rc = png_run_pieces(c, imgs_dev, n, layout, stream, slot, &total, hint, hint_rows);
    if (rc) return rc;
    if (total > out_capacity) {
        c->hint = hint;                      // the caller retries with a larger buffer: the hint is still good for that call
        c->hint_rows = hint_rows;
        return pfail(c, RIG_E_CAPACITY, "the PNG files need %llu bytes, out_capacity is %zu", total, out_capacity);
    }
    if (c->out_caps[slot] < total) {
        if (c->d_outs[slot]) PCK(c, cudaFree(c->d_outs[slot]));
        c->d_outs[slot] = nullptr;
        const size_t want = (size_t)(total + total / 2 + (1u << 20));
        PCK(c, cudaMalloc(&c->d_outs[slot], want));
        c->out_caps[slot] = want;
    }
    rc = png_run_assemble(c, n, c->d_outs[slot], slot);
    if (rc) return rc;
    // the copy runs on its own stream: the next batch's kernels start right behind k_png_assemble
    PCK(c, cudaEventRecord(c->ev_asm, c->stream));
    PCK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_asm, 0));
    {
        TimelineScope tl("d2h_files", c->copy_stream);
        PCK(c, cudaMemcpyAsync(out_host, c->d_outs[slot], total, cudaMemcpyDeviceToHost, c->copy_stream));
    }
    PCK(c, cudaEventRecord(c->ev_copy[slot], c->copy_stream));
    c->slot_busy[slot] = true;
    c->next_slot++;
    return slot;
}

int rig_png_encode_end(rig_png_ctx *c, int slot, uint64_t *offsets_host)
{
    if (!c || slot < 0 || slot > 1 || !offsets_host) return RIG_E_INVALID;
    if (!c->slot_busy[slot]) return pfail(c, RIG_E_INVALID, "slot %d has no batch in flight", slot);
    PCK(c, cudaSetDevice(c->device));
    const int n = c->slot_n[slot];
    if (n) {
        PCK(c, cudaEventSynchronize(c->ev_copy[slot]));
        if (c->profiling) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, c->ev0[slot], c->ev1[slot]) == cudaSuccess) {
                c->stats.encode_ms += ms;
                c->stats.timed_images += (uint64_t)n;
            }
        }
    }
    for (int i = 0; i <= n; i++) offsets_host[i] = c->h_offs[slot][i];
    c->stats.images += (uint64_t)n;
    c->stats.png_bytes += c->h_offs[slot][n];
    c->slot_busy[slot] = false;
    return RIG_OK;
}

int rig_png_encode(rig_png_ctx *c, const uint8_t *imgs_dev, int n, int layout, uint8_t *out, size_t out_capacity, uint64_t *offsets_host,
                   int out_is_device, void *stream)
{
    if (!c) return RIG_E_INVALID;
    if (!offsets_host) return pfail(c, RIG_E_INVALID, "offsets_host is NULL");
    if (!out_is_device) {
        const int slot = rig_png_encode_begin(c, imgs_dev, n, layout, out, out_capacity, stream);
        if (slot < 0) return slot;
        return rig_png_encode_end(c, slot, offsets_host);
    }
    const uint8_t *hint = c->hint;
    const int hint_rows = c->hint_rows;
    c->hint = nullptr;
    int rc = png_check_args(c, imgs_dev, n, layout, out);
    if (rc) return rc;
    offsets_host[0] = 0;
    if (n == 0) return RIG_OK;
    PCK(c, cudaSetDevice(c->device));
    const int slot = (int)(c->next_slot & 1u);
    if (c->slot_busy[slot]) return pfail(c, RIG_E_INVALID, "both output slots are in flight: call rig_png_encode_end first");
    unsigned long long total = 0;
    rc = png_run_pieces(c, imgs_dev, n, layout, stream, slot, &total, hint, hint_rows);
    if (rc) return rc;
    if (total > out_capacity) return pfail(c, RIG_E_CAPACITY, "the PNG files need %llu bytes, out_capacity is %zu", total, out_capacity);
    rc = png_run_assemble(c, n, out, slot);
    if (rc) return rc;
    PCK(c, cudaStreamSynchronize(c->stream));
    if (c->profiling) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->ev0[slot], c->ev1[slot]) == cudaSuccess) {
            c->stats.encode_ms += ms;
            c->stats.timed_images += (uint64_t)n;
        }
    }
    for (int i = 0; i <= n; i++) offsets_host[i] = c->h_offs[slot][i];
    c->stats.images += (uint64_t)n;
    c->stats.png_bytes += total;
    return RIG_OK;
}

int rig_png_set_blank_hint(rig_png_ctx *c, const uint8_t *flags_dev, int rows_per_band)
{
    if (!c) return RIG_E_INVALID;
    if (flags_dev && rows_per_band <= 0) return pfail(c, RIG_E_INVALID, "rows_per_band must be positive");
    c->hint = flags_dev;
    c->hint_rows = rows_per_band;
    return RIG_OK;
}

int rig_png_wait_event(rig_png_ctx *c, void *event)
{
    if (!c || !event) return RIG_E_INVALID;
    PCK(c, cudaSetDevice(c->device));
    PCK(c, cudaStreamWaitEvent(c->stream, (cudaEvent_t)event, 0));
    return RIG_OK;
}

int rig_png_staging(rig_png_ctx *c, int n, uint8_t **out_dev)
{
    if (!c || !out_dev || n < 0) return RIG_E_INVALID;
    PCK(c, cudaSetDevice(c->device));
    const size_t need = (size_t)std::max(n, 1) * c->W * c->H * 3;
    if (c->stage_cap < need) {
        PCK(c, cudaDeviceSynchronize());
        if (c->d_stage) PCK(c, cudaFree(c->d_stage));
        c->d_stage = nullptr;
        PCK(c, cudaMalloc(&c->d_stage, need));
        c->stage_cap = need;
    }
    *out_dev = c->d_stage;
    return RIG_OK;
}

int rig_png_debug_code(int channels, uint8_t *len_out, uint32_t *hdr_out, uint32_t *hdr_bits_out, uint32_t *lut_out)
{
    if (!len_out || !hdr_out || !hdr_bits_out || !lut_out || (channels != 3 && channels != 4)) return RIG_E_INVALID;
    PngCode code;
    if (!png_build_code(channels, code)) return RIG_E_UNSUPPORTED;
    memcpy(len_out, code.len, PNG_NLIT);
    memcpy(hdr_out, code.hdr, sizeof code.hdr);
    *hdr_bits_out = code.hdr_bits;
    for (uint32_t v = 0; v < 256; v++) lut_out[v] = png_lut_literal(code, v);
    lut_out[256] = lut_out[257] = lut_out[258] = 0;
    for (uint32_t l = 3; l <= 258; l++) lut_out[256 + l] = png_lut_match(code, l);
    lut_out[256 + 259] = png_lut_literal(code, 256);
    return RIG_OK;
}

int rig_png_get_stats(rig_png_ctx *c, rig_png_stats *out)
{
    if (!c || !out) return RIG_E_INVALID;
    *out = c->stats;
    return RIG_OK;
}

int rig_png_set_profiling(rig_png_ctx *c, int on)
{
    if (!c) return RIG_E_INVALID;
    c->profiling = on != 0;
    if (on) { c->stats.encode_ms = 0; c->stats.timed_images = 0; }
    return RIG_OK;
}

}  // extern "C"
