// mcg_png_parse / mcg_png_decode (include/mcgaze_b200.h): PNG files -> BGR uint8 frames in HBM, the part of
// LoadImageFromFile (mmdet/datasets/pipelines/loading.py:58-69 -> mmcv.imfrombytes -> cv2.imdecode) that used to stay on
// the host.  The host only walks the chunk list (signature, IHDR, PLTE, IDAT payloads, optional CRC check); inflate and
// scanline reconstruction run on the device, one warp per image (png_core.cuh).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/mcgaze_b200.h"
#include "common.cuh"
#include "png_core.cuh"

namespace mcg {

constexpr int kInflateWarps = 1;   // one image per CTA: the decoding tables sit at a constant shared-memory offset
constexpr int kUnfilterWarps = 4;

__device__ __forceinline__ bool png_job_ok(const mcg_png_job& j) {
  return png::channels_of(j.color_type) != 0 && j.zdata != nullptr && j.zbytes > 0 && j.scan != nullptr && j.dst != nullptr &&
         j.width > 0 && j.height > 0 && j.dst_stride >= 3LL * j.width && (j.color_type != 3 || j.palette != nullptr);
}

// The job table lives in HBM (the caller uploads it with the compressed bytes), so ONE launch covers a whole batch:
// a warp spends ~10 ms on a 300 x 300 photograph whatever else runs, and throughput comes from the thousands of warps a
// B200 keeps resident (148 SMs x 32 one-warp CTAs), not from the speed of one stream.
__global__ void __launch_bounds__(32 * kInflateWarps) png_inflate_kernel(const mcg_png_job* __restrict__ jobs, int n,
                                                                          int32_t* __restrict__ status) {
  __shared__ png::Tables tables[kInflateWarps];
  const int warp = kInflateWarps == 1 ? 0 : threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * kInflateWarps + warp;
  if (idx >= n) return;
  const mcg_png_job job = jobs[idx];
  if (!png_job_ok(job)) {
    if (lane == 0) status[idx] = MCG_PNG_BAD_JOB;
    return;
  }
  const long long expected = static_cast<long long>(job.height) * (1 + static_cast<long long>(job.width) * png::channels_of(job.color_type));
  long long produced = 0;
  int st = png::inflate_warp(job.zdata, job.zbytes, job.scan, expected, tables[warp], lane, &produced);
  if (st == png::ST_OK && produced != expected) st = png::ST_OUTPUT_SHORT;
  if (lane == 0) status[idx] = st;
}

__global__ void __launch_bounds__(32 * kUnfilterWarps) png_unfilter_kernel(const mcg_png_job* __restrict__ jobs, int n,
                                                                            int32_t* __restrict__ status) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * kUnfilterWarps + warp;
  if (idx >= n) return;
  if (status[idx] != png::ST_OK) return;     // uniform per warp: the inflate kernel wrote it
  const mcg_png_job job = jobs[idx];
  const int ct = job.color_type, bpp = png::channels_of(ct);
  const int H = job.height;
  const int rowbytes = job.width * bpp;
  const long long stride = 1 + static_cast<long long>(rowbytes);
  bool bad = false;
  for (int band = 0; band * 32 < H; ++band) {
    const int r = band * 32 + lane;
    const bool valid = r < H;
    uint8_t* row = job.scan + static_cast<long long>(valid ? r : 0) * stride + 1;
    uint8_t* dst_row = job.dst + static_cast<long long>(valid ? r : 0) * job.dst_stride;
    int ft = valid ? row[-1] : 0;
    if (ft > 4) {
      bad = true;
      ft = 0;
    }
    png::LaneState s = {0u, 0u, 0u, 0, 0};
    // Row above the band (lane 0's "up" bytes; lane 31 of the band before left it in `scan`): the warp fetches it 32
    // bytes at a time, one coalesced load 32 steps ahead of use, and hands lane 0 its byte by shuffle - a load per step
    // in lane 0 would put an L2 round trip on every step of the wavefront.
    const uint8_t* above = job.scan + (static_cast<long long>(band) * 32 - 1) * stride + 1;
    uint32_t cur32 = (band > 0 && lane < rowbytes) ? png::load_cg(above + lane) : 0u;
    uint32_t nxt32 = (band > 0 && 32 + lane < rowbytes) ? png::load_cg(above + 32 + lane) : 0u;
    uint32_t raw_next = (valid && lane == 0 && rowbytes > 0) ? row[0] : 0u;    // this row's next filtered byte, one step ahead
    for (int t = 0; t < rowbytes + 31; ++t) {
      if ((t & 31) == 0 && t > 0) {
        cur32 = nxt32;
        nxt32 = (band > 0 && t + 32 + lane < rowbytes) ? png::load_cg(above + t + 32 + lane) : 0u;
      }
      const uint32_t up0 = __shfl_sync(0xffffffffu, cur32, t & 31);
      uint32_t up = __shfl_up_sync(0xffffffffu, s.last, 1);
      if (lane == 0) up = up0;
      const int j = t - lane;
      const bool active = valid && j >= 0 && j < rowbytes;
      const uint32_t raw = raw_next;
      if (valid && j + 1 >= 0 && j + 1 < rowbytes) raw_next = row[j + 1];
      if (active) {
        png::unfilter_byte(s, ft, raw, up, bpp, ct, job.palette, dst_row);
        if (lane == 31) row[j] = static_cast<uint8_t>(s.last);                         // ... for the next band's lane 0
      }
    }
    __syncwarp();
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) status[idx] = png::ST_BAD_FILTER;
}

// ---- host: chunk walk ---------------------------------------------------------------------------------------
namespace {
uint32_t g_crc_table[4][256];
std::once_flag g_crc_once;
void crc_init() {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
    g_crc_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 4; ++t) g_crc_table[t][i] = (g_crc_table[t - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[t - 1][i] & 255u];
}
uint32_t crc32_update(uint32_t crc, const uint8_t* p, int64_t n) {
  crc = ~crc;
  while (n >= 4) {
    crc ^= static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8) | (static_cast<uint32_t>(p[2]) << 16) |
           (static_cast<uint32_t>(p[3]) << 24);
    crc = g_crc_table[3][crc & 255u] ^ g_crc_table[2][(crc >> 8) & 255u] ^ g_crc_table[1][(crc >> 16) & 255u] ^
          g_crc_table[0][crc >> 24];
    p += 4;
    n -= 4;
  }
  while (n-- > 0) crc = g_crc_table[0][(crc ^ *p++) & 255u] ^ (crc >> 8);
  return ~crc;
}
inline uint32_t be32(const uint8_t* p) {
  return (static_cast<uint32_t>(p[0]) << 24) | (static_cast<uint32_t>(p[1]) << 16) | (static_cast<uint32_t>(p[2]) << 8) | p[3];
}

}  // namespace

// -> MCG_OK, or MCG_ERR_INVALID with g_last_error-style message in `why`
int png_parse(const uint8_t* f, int64_t n, int check_crc, mcg_png_info* info, uint8_t* z, int64_t zcap, const char** why) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  *why = "";
  if (!f || !info || n < 8 + 25 || std::memcmp(f, sig, 8) != 0) {
    *why = "not a PNG file (signature)";
    return MCG_ERR_INVALID;
  }
  if (check_crc) std::call_once(g_crc_once, crc_init);
  std::memset(info, 0, sizeof(*info));
  int64_t pos = 8, zpos = 0;
  bool have_ihdr = false, have_end = false, overflow = false;
  while (pos + 12 <= n) {
    const uint32_t len = be32(f + pos);
    const uint8_t* type = f + pos + 4;
    const uint8_t* data = f + pos + 8;
    if (len > 0x7fffffffu || pos + 12 + static_cast<int64_t>(len) > n) {
      *why = "truncated chunk";
      return MCG_ERR_INVALID;
    }
    if (check_crc && crc32_update(0u, type, 4 + static_cast<int64_t>(len)) != be32(data + len)) {
      *why = "chunk CRC mismatch";
      return MCG_ERR_INVALID;
    }
    if (!have_ihdr) {
      if (std::memcmp(type, "IHDR", 4) != 0 || len != 13) {
        *why = "first chunk is not IHDR";
        return MCG_ERR_INVALID;
      }
      info->width = static_cast<int32_t>(be32(data));
      info->height = static_cast<int32_t>(be32(data + 4));
      info->bit_depth = data[8];
      info->color_type = data[9];
      info->interlace = data[12];
      if (info->width <= 0 || info->height <= 0 || data[10] != 0 || data[11] != 0) {
        *why = "bad IHDR";
        return MCG_ERR_INVALID;
      }
      const int ct = info->color_type;
      info->channels = ct == 0 ? 1 : ct == 2 ? 3 : ct == 3 ? 1 : ct == 4 ? 2 : ct == 6 ? 4 : 0;
      if (info->channels == 0) {
        *why = "bad colour type";
        return MCG_ERR_INVALID;
      }
      have_ihdr = true;
    } else if (std::memcmp(type, "PLTE", 4) == 0) {
      if (len % 3 != 0 || len > 768) {
        *why = "bad PLTE";
        return MCG_ERR_INVALID;
      }
      std::memcpy(info->palette, data, len);
      info->has_palette = 1;
    } else if (std::memcmp(type, "IDAT", 4) == 0) {
      if (z != nullptr && zpos + static_cast<int64_t>(len) <= zcap)
        std::memcpy(z + zpos, data, len);
      else if (z != nullptr)
        overflow = true;
      zpos += len;
    } else if (std::memcmp(type, "IEND", 4) == 0) {
      have_end = true;
      break;
    }
    pos += 12 + static_cast<int64_t>(len);
  }
  info->idat_bytes = zpos;
  if (!have_ihdr || !have_end || zpos == 0) {
    *why = "missing IHDR / IDAT / IEND";
    return MCG_ERR_INVALID;
  }
  if (info->color_type == 3 && !info->has_palette) {
    *why = "palette image without PLTE";
    return MCG_ERR_INVALID;
  }
  info->supported = (info->bit_depth == 8 && info->interlace == 0) ? 1 : 0;
  if (overflow) {
    *why = "IDAT buffer too small (see idat_bytes)";
    return MCG_ERR_INVALID;
  }
  return MCG_OK;
}

// File sizes (bytes; -1 = cannot stat): what the caller needs to lay out the staging block.
void png_file_sizes(const char* const* paths, int n, int64_t* sizes) {
  for (int i = 0; i < n; ++i) {
    struct stat sb;
    sizes[i] = (paths[i] && ::stat(paths[i], &sb) == 0 && S_ISREG(sb.st_mode)) ? static_cast<int64_t>(sb.st_size) : -1;
  }
}

// The host side of a batch in one call: every file is read, its chunks are walked and its IDAT payload is copied into
// the caller's (pinned) block at slot_off[i] (slot_cap[i] bytes), on `threads` worker threads.  results[i]: 0 = staged, 1 = unreadable, 2 = not a PNG / malformed, 3 = a PNG the device decoder does not take.
// -> number of files with results != 0.
int png_stage_files(const char* const* paths, int n, int check_crc, int threads, const int64_t* slot_off,
                    const int64_t* slot_cap, uint8_t* block, mcg_png_info* infos, int32_t* results) {
  std::atomic<int> next(0), failed(0);
  auto work = [&]() {
    std::vector<uint8_t> buf;
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n) return;
      results[i] = 1;
      std::memset(&infos[i], 0, sizeof(mcg_png_info));
      const int fd = paths[i] ? ::open(paths[i], O_RDONLY | O_CLOEXEC) : -1;
      struct stat sb;
      if (fd < 0 || ::fstat(fd, &sb) != 0 || sb.st_size <= 0) {
        if (fd >= 0) ::close(fd);
        failed.fetch_add(1);
        continue;
      }
      // read() into a per-thread buffer: one sequential copy out of the page cache (mapping every file costs a page
      // fault per 4 KB and a munmap: measured 2.5x slower for 200 KB frames)
      const size_t size = static_cast<size_t>(sb.st_size);
      if (buf.size() < size) buf.resize(size + (size >> 2));
      size_t got = 0;
      while (got < size) {
        const ssize_t r = ::read(fd, buf.data() + got, size - got);
        if (r <= 0) break;
        got += static_cast<size_t>(r);
      }
      ::close(fd);
      if (got != size) {
        failed.fetch_add(1);
        continue;
      }
      const char* why = "";
      const int rc = png_parse(buf.data(), static_cast<int64_t>(size), check_crc, &infos[i], block + slot_off[i], slot_cap[i], &why);
      results[i] = rc != MCG_OK ? 2 : (infos[i].supported ? 0 : 3);
      if (results[i] != 0) failed.fetch_add(1);
    }
  };
  const int nt = std::max(1, std::min(threads, n));
  if (nt == 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (int t = 0; t < nt; ++t) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  return failed.load();
}

void png_decode_launch(const mcg_png_job* jobs, int n, int32_t* status, cudaStream_t st, int* launches) {
  MCG_CHECK(jobs != nullptr && n > 0 && status != nullptr, "null argument");
  cudaPointerAttributes attr;
  MCG_CHECK(cudaPointerGetAttributes(&attr, jobs) == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged),
            "the job table must be in device memory (upload it with the compressed bytes)");
  png_inflate_kernel<<<static_cast<unsigned>((n + kInflateWarps - 1) / kInflateWarps), 32 * kInflateWarps, 0, st>>>(jobs, n, status);
  MCG_CUDA(cudaGetLastError());
  png_unfilter_kernel<<<static_cast<unsigned>((n + kUnfilterWarps - 1) / kUnfilterWarps), 32 * kUnfilterWarps, 0, st>>>(jobs, n, status);
  MCG_CUDA(cudaGetLastError());
  if (launches) *launches = 2;
}

}  // namespace mcg
