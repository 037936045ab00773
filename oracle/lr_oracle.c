/* CPU oracle (C restatement) for the bicubic LR generator.  TEST INFRASTRUCTURE ONLY:
 * linked/loaded only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 *
 * Restates cv2.resize(hr, (W/4, H/4), interpolation=cv2.INTER_CUBIC) on uint8 HWC images, the call
 * the reference makes at src/data/dataset.py:296 and src/data/prepare_data.py:38 (opencv-python
 * >= 4.8, requirements.txt:13; not vendored under /root/reference).  For an exact /4 ratio the
 * cubic taps are [-3, 19, 19, -3] / 32 on pixels 4x..4x+3 and the result is rounded half-to-even:
 *     u = sum a_i a_j hr[4y+i][4x+j][c] ;  lr = clamp((u + 511 + ((u >> 10) & 1)) >> 10, 0, 255)
 * Pinned bit-for-bit against cv2 4.13.0 outputs in tests/golden/lr_*.npz.
 */
#include <stdint.h>
#include <stddef.h>

/* hr: [n][H][W][C] uint8, lr: [n][H/4][W/4][C] uint8.  Returns 0, or -1 on bad arguments. */
int lr_oracle_u8(const uint8_t* hr, uint8_t* lr, long n, int H, int W, int C) {
  static const int a[4] = {-3, 19, 19, -3};
  if (!hr || !lr || n < 0 || H <= 0 || W <= 0 || C <= 0 || (H & 3) || (W & 3)) return -1;
  const int h = H / 4, w = W / 4;
  for (long b = 0; b < n; ++b) {
    const uint8_t* src = hr + (size_t)b * H * W * C;
    uint8_t* dst = lr + (size_t)b * h * w * C;
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x)
        for (int c = 0; c < C; ++c) {
          int32_t u = 0;
          for (int i = 0; i < 4; ++i) {
            const uint8_t* row = src + ((size_t)(4 * y + i) * W + 4 * x) * C + c;
            int32_t r = a[0] * row[0] + a[1] * row[C] + a[2] * row[2 * C] + a[3] * row[3 * C];
            u += a[i] * r;
          }
          int32_t q = (u + 511 + ((u >> 10) & 1)) >> 10;
          dst[((size_t)y * w + x) * C + c] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
        }
  }
  return 0;
}
