// CPU build of the device PNG decoder's core (mcgaze_b200/csrc/png_core.cuh with MCG_PNG_HOST_SIM) for
// tests/test_png.py: the bit reader, the Huffman table construction and the symbol loop run as a one-lane "warp", the
// scanline wavefront as 32 lanes emulated in lock step (lanes are visited from 31 down to 0 inside a step, so the value
// a lane takes from its upper neighbour is the neighbour's value of the step before - what __shfl_up_sync delivers).
// Test infrastructure only; the product path is the CUDA build of the same header (png_decode.cu).
#define MCG_PNG_HOST_SIM 1
#include "../mcgaze_b200/csrc/png_core.cuh"

using namespace mcg::png;

extern "C" int sim_inflate(const uint8_t* in, long long n, uint8_t* out, long long cap, long long* produced) {
  static Tables T;
  return inflate_warp(in, n, out, cap, T, 0, produced);
}

extern "C" long long sim_slow_symbols() { return g_slow_symbols; }

extern "C" int sim_unfilter(uint8_t* scan, int W, int H, int color_type, const uint8_t* palette, uint8_t* dst, long long dst_stride) {
  const int bpp = channels_of(color_type);
  if (bpp == 0) return ST_BAD_JOB;
  const int rowbytes = W * bpp;
  const long long stride = 1 + static_cast<long long>(rowbytes);
  bool bad = false;
  for (int band = 0; band * 32 < H; ++band) {
    LaneState s[32];
    int ft[32];
    for (int k = 0; k < 32; ++k) {
      s[k] = LaneState{0u, 0u, 0u, 0, 0};
      const int r = band * 32 + k;
      ft[k] = r < H ? scan[r * stride] : 0;
      if (ft[k] > 4) {
        bad = true;
        ft[k] = 0;
      }
    }
    for (int t = 0; t < rowbytes + 31; ++t) {
      for (int k = 31; k >= 0; --k) {
        const int r = band * 32 + k;
        const int j = t - k;
        const bool active = r < H && j >= 0 && j < rowbytes;
        uint32_t up = k > 0 ? s[k - 1].last : 0u;
        uint8_t* row = scan + r * stride + 1;
        if (k == 0) up = (active && r > 0) ? row[-stride + j] : 0u;
        if (active) {
          unfilter_byte(s[k], ft[k], row[j], up, bpp, color_type, palette, dst + r * dst_stride);
          if (k == 31) row[j] = static_cast<uint8_t>(s[k].last);
        }
      }
    }
  }
  return bad ? ST_BAD_FILTER : ST_OK;
}
