// orb_ingest.cu -- frame ingest: PNG files -> staging area -> the wave pipeline of orb_api.cu.
// Host decode (orb_png.cpp on a pool of threads) and device decode (k_inflate + k_unfilter, orb_ingest_kernels.cuh).
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "orb_internal.h"
#include "orb_ingest_kernels.cuh"
#include "orb_png.h"

#define g_create_error g_orb_create_error
#define get_plan orb_internal_get_plan
#define run_batch orb_internal_run_batch

namespace {

// whole file into `buf`; returns false with errno-style message
bool read_file(const char* path, std::vector<uint8_t>* buf, std::string* err) {
  const int fd = open(path, O_RDONLY);
  if (fd < 0) { *err = std::string("cannot open ") + path; return false; }
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size <= 0) { close(fd); *err = std::string("cannot stat ") + path; return false; }
  buf->resize((size_t)st.st_size);
  size_t got = 0;
  while (got < buf->size()) {
    const ssize_t r = read(fd, buf->data() + got, buf->size() - got);
    if (r <= 0) break;
    got += (size_t)r;
  }
  close(fd);
  if (got != buf->size()) { *err = std::string("short read on ") + path; return false; }
  return true;
}

// Host-decode source: a pool of threads decodes the files in order straight into the pinned area; stage() waits for
// the frames of its wave and queues their copy.
struct HostDecodeSource : WaveSource {
  const char* const* paths; int n, w, h;
  uint8_t* area; size_t slot; int pitch;
  std::vector<std::thread> pool;
  std::atomic<int> next{0};
  std::atomic<bool> stop{false};
  std::mutex mu; std::condition_variable cv;
  std::vector<uint8_t> done;
  int ready = 0;                 // frames [0, ready) are decoded (guarded by mu)
  int err_code = 0; std::string err_msg;

  void work() {
    std::vector<uint8_t> file;
    orbpng::Scratch scratch;
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n || stop.load()) return;
      std::string msg;
      int code = 0;
      if (!read_file(paths[i], &file, &msg)) code = ORB_E_IO;
      else {
        const char* e = orbpng::decode_gray8(file.data(), file.size(), area + (size_t)i * slot, (size_t)pitch, w, h, &scratch);
        if (e) { code = ORB_E_FORMAT; msg = std::string(paths[i]) + ": " + e; }
      }
      std::lock_guard<std::mutex> lk(mu);
      if (code && !err_code) { err_code = code; err_msg = msg; stop.store(true); }
      done[i] = 1;
      cv.notify_all();
    }
  }
  void start(int n_threads) {
    done.assign(n, 0);
    for (int t = 0; t < n_threads; t++) pool.emplace_back([this] { work(); });
  }
  int stage(orb_ctx* ctx, int, int c0, int nc, cudaEvent_t ready_ev) override {
    {
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] {
        while (ready < n && done[ready]) ready++;
        return err_code != 0 || ready >= c0 + nc;
      });
      if (err_code) return fail(ctx, err_code, "%s", err_msg.c_str());
    }
    CK(cudaMemcpyAsync(ctx->d_frames + (size_t)c0 * slot, area + (size_t)c0 * slot, slot * nc, cudaMemcpyHostToDevice, ctx->s_h2d));
    CK(cudaEventRecord(ready_ev, ctx->s_h2d));
    return ORB_OK;
  }
  ~HostDecodeSource() override {
    stop.store(true);
    for (auto& t : pool) t.join();
  }
};

const char* inflate_status_text(int st) {
  switch (st) {
    case orbk::INF_CORRUPT: return "corrupt deflate data";
    case orbk::INF_SIZE: return "image data does not match the frame size";
    case orbk::INF_TRUNCATED: return "truncated deflate data";
    case orbk::INF_TABLE: return "Huffman table larger than the device decoder holds";
    case orbk::INF_FILTER: return "unknown PNG filter type";
    case orbk::INF_CHECKSUM: return "zlib: incorrect data check";
    case orbk::INF_CRC: return "png: chunk CRC mismatch";
    default: return "unknown decode error";
  }
}

// (re)allocates the device-decode areas for n frames of w x h
int ensure_device_decode(orb_ctx* ctx, int n, int w, int h) {
  const size_t raw = ((size_t)(w + 1) * h + orbk::UNF_LEAD + 48 + 15) / 16 * 16;
  const size_t comp = (raw + raw / 64 + 1024 + 511) / 512 * 512 + 512;
  if (n <= ctx->ingest_cap && raw <= ctx->raw_slot && comp <= ctx->comp_slot) return ORB_OK;
  if (ctx->h_comp) cudaFreeHost(ctx->h_comp);
  if (ctx->h_descs) cudaFreeHost(ctx->h_descs);
  if (ctx->h_inf_status) cudaFreeHost(ctx->h_inf_status);
  cudaFree(ctx->d_comp); cudaFree(ctx->d_raw); cudaFree(ctx->d_descs); cudaFree(ctx->d_inf_status); cudaFree(ctx->d_adler);
  ctx->d_adler = nullptr;
  if (ctx->h_crc) cudaFreeHost(ctx->h_crc);
  if (ctx->h_crc_n) cudaFreeHost(ctx->h_crc_n);
  cudaFree(ctx->d_crc); cudaFree(ctx->d_crc_n);
  ctx->h_crc = ctx->d_crc = nullptr; ctx->h_crc_n = ctx->d_crc_n = nullptr;
  ctx->h_comp = nullptr; ctx->h_descs = nullptr; ctx->h_inf_status = nullptr;
  ctx->d_comp = ctx->d_raw = nullptr; ctx->d_descs = nullptr; ctx->d_inf_status = nullptr;
  ctx->ingest_cap = 0;
  const int cap = std::max(n, ctx->p.max_batch);
  CK(cudaHostAlloc((void**)&ctx->h_comp, comp * cap, cudaHostAllocDefault));
  CK(cudaHostAlloc((void**)&ctx->h_descs, sizeof(orbk::InflateDesc) * cap, cudaHostAllocDefault));
  CK(cudaHostAlloc((void**)&ctx->h_inf_status, sizeof(int) * cap, cudaHostAllocDefault));
  CK(cudaMalloc((void**)&ctx->d_comp, comp * cap));
  CK(cudaMalloc((void**)&ctx->d_raw, raw * cap));
  CK(cudaMalloc((void**)&ctx->d_descs, sizeof(orbk::InflateDesc) * cap));
  CK(cudaMalloc((void**)&ctx->d_inf_status, sizeof(int) * cap));
  CK(cudaMalloc((void**)&ctx->d_adler, sizeof(uint32_t) * cap));
  CK(cudaHostAlloc((void**)&ctx->h_crc, sizeof(uint32_t) * 3 * orbk::PNG_CRC_CAP * cap, cudaHostAllocDefault));
  CK(cudaHostAlloc((void**)&ctx->h_crc_n, sizeof(int) * cap, cudaHostAllocDefault));
  CK(cudaMalloc((void**)&ctx->d_crc, sizeof(uint32_t) * 3 * orbk::PNG_CRC_CAP * cap));
  CK(cudaMalloc((void**)&ctx->d_crc_n, sizeof(int) * cap));
  CK(cudaMemset(ctx->d_comp, 0, comp * cap));
  // the inflate kernels are latency chains (one busy lane per warp): give their blocks the first free slots
  int prio_least = 0, prio_greatest = 0;
  CK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  for (cudaStream_t& q : ctx->s_ingest) if (!q) CK(cudaStreamCreateWithPriority(&q, cudaStreamNonBlocking, prio_greatest));
  CK(cudaFuncSetAttribute(orbk::k_inflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(orbk::InflateShared)));
  ctx->comp_slot = comp; ctx->raw_slot = raw; ctx->ingest_cap = cap;
  return ORB_OK;
}

// Device-decode source: host threads only read the files, check the framing and lay the deflate streams out in the pinned
// area; stage() uploads the compressed bytes of a wave and queues k_inflate + k_unfilter on the copy stream.
struct DeviceDecodeSource : WaveSource {
  const char* const* paths; int n, w, h;
  orb_ctx* ctx;
  std::vector<std::thread> pool;
  std::atomic<int> next{0};
  std::atomic<bool> stop{false};
  std::mutex mu; std::condition_variable cv;
  std::vector<uint8_t> done;
  std::vector<uint32_t> bytes;
  int ready = 0;
  int err_code = 0; std::string err_msg;

  void work() {
    std::vector<uint8_t> file;
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n || stop.load()) return;
      std::string msg;
      int code = 0;
      if (!read_file(paths[i], &file, &msg)) code = ORB_E_IO;
      else {
        orbpng::Info I;
        size_t nb = 0;
        // the payload goes to slot + 14: the zlib header ends at + 16, where the (aligned) deflate data starts; the chunk
        // checksums are listed for k_png_crc instead of being computed here (0.23 ms per frame on one core)
        static_assert(sizeof(orbpng::ChunkCrc) == 12, "three words per chunk descriptor");
        int ncrc = 0;
        const char* e = orbpng::extract_deflate(file.data(), file.size(), &I, ctx->h_comp + (size_t)i * ctx->comp_slot + orbk::PNG_PAYLOAD_OFS,
                                                ctx->comp_slot - 512 - orbk::PNG_PAYLOAD_OFS, &nb,
                                                reinterpret_cast<orbpng::ChunkCrc*>(ctx->h_crc + (size_t)i * 3 * orbk::PNG_CRC_CAP),
                                                orbk::PNG_CRC_CAP, &ncrc);
        ctx->h_crc_n[i] = ncrc;
        if (!e && (I.width != w || I.height != h)) e = "png: image size differs from the expected frame size";
        if (!e && (I.color_type != 0 || I.bit_depth != 8)) e = "png: the device decoder takes 8-bit gray files only";
        if (e) { code = ORB_E_FORMAT; msg = std::string(paths[i]) + ": " + e; }
        bytes[i] = (uint32_t)nb;
      }
      std::lock_guard<std::mutex> lk(mu);
      if (code && !err_code) { err_code = code; err_msg = msg; stop.store(true); }
      done[i] = 1;
      cv.notify_all();
    }
  }
  void start(int n_threads) {
    done.assign(n, 0);
    bytes.assign(n, 0);
    for (int t = 0; t < n_threads; t++) pool.emplace_back([this] { work(); });
  }
  // Inflate waves are independent of the ORB waves: a deflate stream is serial (one warp, ~15 ms per KITTI frame), so
  // the decoder wants as many streams in flight as the host has read, while the ORB kernels want short waves that
  // start early.  Inflate wave k covers frames [k * IW, (k + 1) * IW) and runs on ingest stream k % N_INGEST, so
  // consecutive inflate waves overlap; an ORB wave waits for the inflate wave that holds its last frame.
  static constexpr int IW = 256;
  int preferred_wave(const orb_ctx* c) const override { return c->chunk; }   // frames are resident before the ORB waves start
  int inflated = 0, inflate_waves = 0;
  cudaStream_t last_q = nullptr;
  int launch_inflate(orb_ctx* ctx, int f0, int nf) {
    cudaStream_t q = ctx->s_ingest[inflate_waves % orb_ctx::N_INGEST];
    if (inflate_waves < orb_ctx::N_INGEST) CK(cudaStreamWaitEvent(q, ctx->ev_start, 0));
    const uint32_t out_bytes = (uint32_t)((size_t)(w + 1) * h);
    // one strided copy for the wave: every slot up to the longest stream in it (stream + Adler trailer + zero pad)
    size_t up = 0;
    for (int i = f0; i < f0 + nf; i++) up = std::max(up, ((size_t)bytes[i] + 16 + 4 + 16 + 15) / 16 * 16);
    CK(cudaMemcpy2DAsync(ctx->d_comp + (size_t)f0 * ctx->comp_slot, ctx->comp_slot, ctx->h_comp + (size_t)f0 * ctx->comp_slot,
                         ctx->comp_slot, up, nf, cudaMemcpyHostToDevice, q));
    for (int i = f0; i < f0 + nf; i++) {
      ctx->h_descs[i] = orbk::InflateDesc{ctx->d_comp + (size_t)i * ctx->comp_slot + 16, bytes[i], out_bytes, ctx->d_raw + (size_t)i * ctx->raw_slot + orbk::UNF_LEAD};
    }
    CK(cudaMemcpyAsync(ctx->d_descs + f0, ctx->h_descs + f0, sizeof(orbk::InflateDesc) * nf, cudaMemcpyHostToDevice, q));
    orbk::k_inflate<<<nf, 32, sizeof(orbk::InflateShared), q>>>(ctx->d_descs + f0, ctx->d_inf_status + f0, ctx->d_adler + f0);
    orbk::k_unfilter<<<(nf + orbk::UNF_WARPS - 1) / orbk::UNF_WARPS, orbk::UNF_WARPS * 32, 0, q>>>(
        ctx->d_raw + (size_t)f0 * ctx->raw_slot, ctx->raw_slot, ctx->d_frames + (size_t)f0 * ctx->frames_slot_bytes,
        ctx->frames_slot_bytes, ctx->frames_pitch, w, h, nf, ctx->d_inf_status + f0, ctx->d_adler + f0);
    CK(cudaMemcpyAsync(ctx->d_crc + (size_t)f0 * 3 * orbk::PNG_CRC_CAP, ctx->h_crc + (size_t)f0 * 3 * orbk::PNG_CRC_CAP,
                       sizeof(uint32_t) * 3 * orbk::PNG_CRC_CAP * nf, cudaMemcpyHostToDevice, q));
    CK(cudaMemcpyAsync(ctx->d_crc_n + f0, ctx->h_crc_n + f0, sizeof(int) * nf, cudaMemcpyHostToDevice, q));
    orbk::k_png_crc<<<nf, orbk::PNG_CRC_CAP, 0, q>>>(ctx->d_comp + (size_t)f0 * ctx->comp_slot, ctx->comp_slot,
                                                     ctx->d_crc + (size_t)f0 * 3 * orbk::PNG_CRC_CAP, ctx->d_crc_n + f0, ctx->d_inf_status + f0);
    CK(cudaGetLastError());
    ctx->launches += 3;
    inflate_waves++;
    last_q = q;
    return ORB_OK;
  }
  int stage(orb_ctx* ctx, int, int c0, int nc, cudaEvent_t ready_ev) override {
    cudaStream_t also[8];
    int n_also = 0;
    // Measured on B200: ORB kernels and inflate kernels sharing the SMs slow each other far more than the overlap gains
    // (126 ms vs 69 ms per 1024 KITTI frames), so every frame is inflated -- in waves that follow the host reads -- before
    // the first ORB wave starts.  ORB_INGEST_SEPARATE=0 restores the interleaved schedule for experiments.
    static const bool separate = !(getenv("ORB_INGEST_SEPARATE") && atoi(getenv("ORB_INGEST_SEPARATE")) == 0);
    const int want = separate ? n : c0 + nc;
    while (inflated < want) {
      if (last_q && n_also < 8 && inflated > c0) also[n_also++] = last_q;     // this ORB wave spans several inflate waves
      const int f0 = inflated, nf = std::min(IW, n - f0);
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] {
          while (ready < n && done[ready]) ready++;
          return err_code != 0 || ready >= f0 + nf;
        });
        if (err_code) return fail(ctx, err_code, "%s", err_msg.c_str());
      }
      const int rc = launch_inflate(ctx, f0, nf);
      if (rc) return rc;
      inflated = f0 + nf;
    }
    // frames of this ORB wave come from the inflate wave queued last at the latest (earlier ORB waves have waited for
    // the earlier inflate waves, except for those first needed by this very wave)
    for (int k = 0; k < n_also; k++) {
      if (also[k] == last_q) continue;
      CK(cudaEventRecord(ctx->ev_chain, also[k]));
      CK(cudaStreamWaitEvent(last_q, ctx->ev_chain, 0));
    }
    CK(cudaEventRecord(ready_ev, last_q));
    return ORB_OK;
  }
  ~DeviceDecodeSource() override {
    stop.store(true);
    for (auto& t : pool) t.join();
  }
};

int png_status(const char* e) {
  if (!e) return ORB_OK;
  snprintf(g_create_error, sizeof(g_create_error), "%s", e);
  return ORB_E_FORMAT;
}

}  // namespace

extern "C" {

int orb_png_info(const uint8_t* file, size_t file_bytes, orb_image_info* info) {
  if (!file || !info) return ORB_E_INVALID;
  orbpng::Info I;
  const int rc = png_status(orbpng::read_info(file, file_bytes, &I));
  if (rc) return rc;
  info->width = I.width; info->height = I.height; info->bit_depth = I.bit_depth; info->channels = I.channels;
  return ORB_OK;
}

int orb_png_decode_gray8(const uint8_t* file, size_t file_bytes, uint8_t* dst, size_t pitch, int w, int h) {
  if (!file || !dst) return ORB_E_INVALID;
  return png_status(orbpng::decode_gray8(file, file_bytes, dst, pitch, w, h, nullptr));
}

int orb_imread_gray8(const char* path, uint8_t* dst, size_t pitch, int cap_w, int cap_h, int* w, int* h) {
  if (!path || !dst) return ORB_E_INVALID;
  std::vector<uint8_t> file;
  std::string msg;
  if (!read_file(path, &file, &msg)) { snprintf(g_create_error, sizeof(g_create_error), "%s", msg.c_str()); return ORB_E_IO; }
  orbpng::Info I;
  int rc = png_status(orbpng::read_info(file.data(), file.size(), &I));
  if (rc) return rc;
  if (w) *w = I.width;
  if (h) *h = I.height;
  if (I.width > cap_w || I.height > cap_h) { snprintf(g_create_error, sizeof(g_create_error), "%s: image %dx%d exceeds the buffer", path, I.width, I.height); return ORB_E_CAPACITY; }
  return png_status(orbpng::decode_gray8(file.data(), file.size(), dst, pitch, I.width, I.height, nullptr));
}

int orb_detect_and_compute_files(orb_ctx* ctx, const char* const* paths, int n_frames, int n_threads, int decode_on_device,
                                 int cap, orb_keypoint* kps, float* angles, orb_descriptor* desc, int* n_out,
                                 int outputs_on_device) {
  if (!ctx) return ORB_E_INVALID;
  if (!paths || n_frames < 1) return fail(ctx, ORB_E_INVALID, "no frame files");
  if (n_frames > ctx->p.max_batch) return fail(ctx, ORB_E_CAPACITY, "batch %d exceeds max_batch %d", n_frames, ctx->p.max_batch);
  CK(cudaSetDevice(ctx->p.device));
  // the first file fixes the frame size of the call
  int w = 0, h = 0;
  {
    uint8_t head[64];
    const int fd = open(paths[0], O_RDONLY);
    const ssize_t got = fd >= 0 ? read(fd, head, sizeof(head)) : -1;
    if (fd >= 0) close(fd);
    if (got < 33) return fail(ctx, ORB_E_IO, "cannot read %s", paths[0]);
    orbpng::Info I;
    const char* e = orbpng::read_info(head, (size_t)got, &I);
    if (e) return fail(ctx, ORB_E_FORMAT, "%s: %s", paths[0], e);
    w = I.width; h = I.height;
  }
  int rc = get_plan(ctx, w, h);
  if (rc) return rc;
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  n_threads = std::min(n_threads, n_frames);
  if (decode_on_device) {
    if ((rc = ensure_device_decode(ctx, n_frames, w, h))) return rc;
    DeviceDecodeSource src;
    src.paths = paths; src.n = n_frames; src.w = w; src.h = h; src.ctx = ctx;
    src.start(n_threads);
    rc = run_batch(ctx, nullptr, 0, n_frames, w, h, 0, 0, cap, kps, angles, desc, n_out, outputs_on_device, &src);
    if (rc) return rc;
    // decode failures are per-frame flags on the device: the results of a failed frame are meaningless, report it
    for (cudaStream_t q : ctx->s_ingest) CK(cudaStreamSynchronize(q));
    CK(cudaMemcpy(ctx->h_inf_status, ctx->d_inf_status, sizeof(int) * n_frames, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n_frames; i++)
      if (ctx->h_inf_status[i]) return fail(ctx, ORB_E_FORMAT, "%s: %s", paths[i], inflate_status_text(ctx->h_inf_status[i]));
    return ORB_OK;
  }
  const size_t need = ctx->frames_slot_bytes * (size_t)n_frames;
  if (need > ctx->h_ingest_bytes) {
    if (ctx->h_ingest) CK(cudaFreeHost(ctx->h_ingest));
    ctx->h_ingest = nullptr; ctx->h_ingest_bytes = 0;
    CK(cudaHostAlloc((void**)&ctx->h_ingest, need, cudaHostAllocDefault));
    ctx->h_ingest_bytes = need;
    memset(ctx->h_ingest, 0, need);
  }
  HostDecodeSource src;
  src.paths = paths; src.n = n_frames; src.w = w; src.h = h;
  src.area = ctx->h_ingest; src.slot = ctx->frames_slot_bytes; src.pitch = ctx->frames_pitch;
  src.start(n_threads);
  return run_batch(ctx, nullptr, 0, n_frames, w, h, 0, 0, cap, kps, angles, desc, n_out, outputs_on_device, &src);
}

int orb_get_ingested_frame(orb_ctx* ctx, int frame, uint8_t* dst, size_t dst_pitch, int* w, int* h) {
  if (!ctx || !dst) return ORB_E_INVALID;
  if (!ctx->plan_valid || ctx->last_frames != ctx->d_frames || frame < 0 || frame >= ctx->last_n)
    return fail(ctx, ORB_E_INVALID, "frame %d is not in the staging area", frame);
  const int W = ctx->plan.lv[0].w, H = ctx->plan.lv[0].h;
  if (dst_pitch < (size_t)W) return fail(ctx, ORB_E_INVALID, "dst_pitch too small");
  CK(cudaSetDevice(ctx->p.device));
  CK(cudaMemcpy2DAsync(dst, dst_pitch, ctx->d_frames + (size_t)frame * ctx->frames_slot_bytes, ctx->frames_pitch, W, H,
                       cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (w) *w = W;
  if (h) *h = H;
  return ORB_OK;
}


int orb_debug_inflate(orb_ctx* ctx, const uint8_t* streams, const uint32_t* offsets, int n, uint8_t* out,
                      const uint32_t* out_offsets, int* status) {
  if (!ctx || !streams || !offsets || !out || !out_offsets || !status || n < 1) return ORB_E_INVALID;
  CK(cudaSetDevice(ctx->p.device));
  // every stream in its own zero-padded, 512-byte aligned slot; outputs 16-byte aligned
  std::vector<size_t> in_at(n + 1, 0), out_at(n + 1, 0);
  for (int i = 0; i < n; i++) {
    in_at[i + 1] = in_at[i] + ((size_t)(offsets[i + 1] - offsets[i]) + 16 + 511) / 512 * 512 + 512;
    out_at[i + 1] = out_at[i] + ((size_t)(out_offsets[i + 1] - out_offsets[i]) + 32 + 15) / 16 * 16;
  }
  std::vector<uint8_t> packed(in_at[n], 0);
  for (int i = 0; i < n; i++) memcpy(packed.data() + in_at[i], streams + offsets[i], offsets[i + 1] - offsets[i]);
  uint8_t *d_in = nullptr, *d_out = nullptr;
  orbk::InflateDesc* d_desc = nullptr;
  int* d_st = nullptr;
  CK(cudaMalloc((void**)&d_in, in_at[n]));
  CK(cudaMalloc((void**)&d_out, out_at[n]));
  CK(cudaMalloc((void**)&d_desc, sizeof(orbk::InflateDesc) * n));
  CK(cudaMalloc((void**)&d_st, sizeof(int) * n));
  std::vector<orbk::InflateDesc> descs(n);
  for (int i = 0; i < n; i++)
    descs[i] = orbk::InflateDesc{d_in + in_at[i], offsets[i + 1] - offsets[i], out_offsets[i + 1] - out_offsets[i], d_out + out_at[i]};
  CK(cudaMemcpy(d_in, packed.data(), in_at[n], cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_desc, descs.data(), sizeof(orbk::InflateDesc) * n, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out, 0xEE, out_at[n]));
  CK(cudaFuncSetAttribute(orbk::k_inflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(orbk::InflateShared)));
  orbk::k_inflate<<<n, 32, sizeof(orbk::InflateShared), ctx->stream>>>(d_desc, d_st, nullptr);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(status, d_st, sizeof(int) * n, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++)
    CK(cudaMemcpy(out + out_offsets[i], d_out + out_at[i], out_offsets[i + 1] - out_offsets[i], cudaMemcpyDeviceToHost));
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_desc); cudaFree(d_st);
  return ORB_OK;
}

}  // extern "C"
