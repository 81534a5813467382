// bm.cu — cv::StereoBM::compute for sm_100a (BASELINE config 2; the reference never calls StereoBM,
// grep: no "StereoBM" under /root/reference).  Arithmetic: SURVEY.md Appendix A.9, XSOBEL pre-filter, no LR check.
// minDisparity != 0: disparity index k stands for d = k + minDisparity, i.e. the right image is read minDisparity
// columns further left; output is produced where every window pixel has all D partners inside the image
// (cv::getValidDisparityROI's x range) and is FILTERED elsewhere.  Structure mirrors the SGBM cost kernel (sgbm.cu, K2): a CTA walks
// down a band of rows keeping |PL-PR| of the last blockSize+1 rows in a shared-memory ring and running
// vertical sums; the horizontal window gives the SAD of 8 disparities per lane, and the lane group of
// a pixel selects the winner with CREDUX without ever writing a cost volume to HBM.
#include "common.cuh"

namespace svgpu {

__global__ void __launch_bounds__(256) k_bm_prefilter(const uint8_t* __restrict__ left, size_t lstride,
                                                      const uint8_t* __restrict__ right, size_t rstride,
                                                      uint8_t* __restrict__ pfL, uint8_t* __restrict__ pfR, int W, int H, int cap) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const uint8_t* img = blockIdx.z ? right : left;
  const size_t stride = blockIdx.z ? rstride : lstride;
  uint8_t* out = (blockIdx.z ? pfR : pfL) + (size_t)y * W + x;
  const int He = H & ~1;  // rows are produced in pairs; an unpaired last row is all `cap`
  if (y >= He || x == 0 || x == W - 1) { *out = (uint8_t)cap; return; }
  const int yu = y > 0 ? y - 1 : (H > 1 ? 1 : 0), yd = y < H - 1 ? y + 1 : (H > 1 ? H - 2 : 0);
  const uint8_t* r = img + (size_t)y * stride;
  const uint8_t* u = img + (size_t)yu * stride;
  const uint8_t* d = img + (size_t)yd * stride;
  const int v = ((int)u[x + 1] - (int)u[x - 1]) + 2 * ((int)r[x + 1] - (int)r[x - 1]) + ((int)d[x + 1] - (int)d[x - 1]);
  *out = (uint8_t)(min(max(v, -cap), cap) + cap);
}

static constexpr int kBmTW = 32;
static constexpr int kBmThreads = 256;

struct BmSmem {
  int NC, NR, NRp, RS;
  size_t offLeft, offRA, offRB, offVs, offTex, offRing, total;
};
__host__ __device__ inline BmSmem bm_smem_layout(int D, int r) {
  BmSmem s;
  s.NC = kBmTW + 2 * r;
  s.NR = s.NC + D - 1;
  s.NRp = (s.NR + 3) & ~1;
  s.RS = 2 * r + 2;
  size_t o = 0;
  s.offLeft = o; o += (size_t)s.NC * 4;
  s.offRA = o; o += (size_t)s.NRp * 2;
  s.offRB = o; o += (size_t)s.NRp * 2;
  s.offTex = o; o += (size_t)s.NC * 4 + (size_t)s.RS * s.NC;  // vertical texture sums + ring of |PL-cap|
  o = (o + 15) & ~(size_t)15;
  s.offVs = o; o += (size_t)s.NC * D * 2;
  s.offRing = o; o += (size_t)s.RS * s.NC * D;
  s.total = o;
  return s;
}

template <int LPP>
__global__ void __launch_bounds__(kBmThreads) k_bm_sad(const uint8_t* __restrict__ pfL, const uint8_t* __restrict__ pfR,
                                                       int16_t* __restrict__ disp, size_t dstride, BmCfg cfg, int rowsPerBand) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int D = cfg.D, r = cfg.bs / 2, W = cfg.W, H = cfg.H;
  const int minX1 = max(D - 1 + cfg.minD, 0), W1 = W - minX1;  // computed columns x' = x - minX1
  const int xpEnd = W1 - r + min(cfg.minD, 0);                  // valid ROI in x' is [r, xpEnd)
  const BmSmem L = bm_smem_layout(D, r);
  uint32_t* sLeft = (uint32_t*)(smem + L.offLeft);
  uint16_t* sRA = (uint16_t*)(smem + L.offRA);
  uint16_t* sRB = (uint16_t*)(smem + L.offRB);
  int* sTexV = (int*)(smem + L.offTex);
  uint8_t* sTexRing = smem + L.offTex + (size_t)L.NC * 4;
  uint16_t* sVs = (uint16_t*)(smem + L.offVs);
  uint8_t* sRing = smem + L.offRing;
  const int NC = L.NC, NR = L.NR, RS = L.RS;

  const int x0 = blockIdx.x * kBmTW;
  // only rows [r, H-r) produce output; a band covers output rows [yb, ye)
  const int yb = r + blockIdx.y * rowsPerBand, ye = min(yb + rowsPerBand, H - r);
  if (yb >= ye) return;
  const int lane = threadIdx.x & 31, gl = lane % LPP, d0 = gl * 8;
  const unsigned gmask = (LPP == 32) ? 0xFFFFFFFFu : (((1u << LPP) - 1u) << (lane - gl));
  const bool active = d0 < D;
  const int nslot = kBmThreads / LPP, slot = threadIdx.x / LPP;
  const int xrMax = x0 + kBmTW - 1 + r + minX1 - cfg.minD;

  for (int i = threadIdx.x; i < NC * (D >> 3); i += kBmThreads) ((uint4*)sVs)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < NC; i += kBmThreads) sTexV[i] = 0;

  // rows yb-r .. ye-1+r enter one at a time; output row y is complete once row y+r has been added
  for (int yr = yb - r; yr < ye + r; yr++) {
    __syncthreads();
    const uint8_t* rowL = pfL + (size_t)yr * W;
    const uint8_t* rowR = pfR + (size_t)yr * W;
    const int subRow = yr - (2 * r + 1);  // leaves the window when yr enters (if it was ever added)
    for (int c = threadIdx.x; c < NC; c += kBmThreads) {
      const int xp = x0 - r + c;
      uint32_t v = 0;
      int t = 0;
      if (xp >= 0 && xp < W1) { v = rowL[xp + minX1]; t = abs((int)v - cfg.cap); }
      sLeft[c] = v | (v << 16);
      int tv = sTexV[c] + t;
      if (subRow >= yb - r) tv -= sTexRing[(subRow % RS) * NC + c];
      sTexV[c] = tv;
      sTexRing[(yr % RS) * NC + c] = (uint8_t)t;
    }
    for (int j = threadIdx.x; j < NR; j += kBmThreads) {
      const int xr = xrMax - j;
      const uint16_t v = (xr >= 0 && xr < W) ? rowR[xr] : 0;
      sRA[j] = v;
      if (j > 0) sRB[j - 1] = v;
    }
    __syncthreads();
    if (active && slot < nslot) {
      uint8_t* ring = sRing + (size_t)(yr % RS) * NC * D;
      const uint8_t* rs = subRow >= yb - r ? sRing + (size_t)(subRow % RS) * NC * D : nullptr;
      for (int c = slot; c < NC; c += nslot) {
        const int xp = x0 - r + c;
        if (xp < 0 || xp >= W1) continue;
        const uint32_t u = sLeft[c];
        const int j0 = (NC - 1 - c) + d0;
        const uint32_t* rp = (const uint32_t*)((j0 & 1) ? sRB : sRA) + (j0 >> 1);
        uint32_t pix[4];
#pragma unroll
        for (int q = 0; q < 4; q++) pix[q] = __vabsdiffu2(u, rp[q]);
        uint2 packed;
        packed.x = __byte_perm(pix[0], pix[1], 0x6420);
        packed.y = __byte_perm(pix[2], pix[3], 0x6420);
        *(uint2*)(ring + (size_t)c * D + d0) = packed;
        uint4* vp = (uint4*)(sVs + (size_t)c * D + d0);
        uint4 v = *vp;
        v.x += pix[0]; v.y += pix[1]; v.z += pix[2]; v.w += pix[3];
        if (rs) {
          const uint2 s = *(const uint2*)(rs + (size_t)c * D + d0);
          v.x -= __byte_perm(s.x, 0, 0x4140); v.y -= __byte_perm(s.x, 0, 0x4342);
          v.z -= __byte_perm(s.y, 0, 0x4140); v.w -= __byte_perm(s.y, 0, 0x4342);
        }
        *vp = v;
      }
    }
    const int y = yr - r;  // output row completed by this row
    if (y < yb) continue;
    __syncthreads();
    // warp-uniform loop over the tile's output pixels so that ballots stay converged
    for (int t0 = 0; t0 < kBmTW; t0 += nslot) {
      const int t = t0 + slot;
      const int xp = x0 + t, x = xp + minX1;
      const bool on = active && slot < nslot && t < kBmTW && xp >= r && xp < xpEnd;  // valid ROI in x (A.9)
      uint32_t sad[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
      int tex = 0;
      if (on) {
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (int dx = -r; dx <= r; dx++) {
          const int cc = xp + dx - (x0 - r);
          const uint4 v = *(const uint4*)(sVs + (size_t)cc * D + d0);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          tex += sTexV[cc];
        }
        sad[0] = acc.x; sad[1] = acc.y; sad[2] = acc.z; sad[3] = acc.w;
      }
      // winner: first strict minimum scanning d from D-1 down to 0, i.e. min over (sad, k = D-1-d)
      uint32_t key = 0xFFFFFFFFu;
      if (on) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          key = min(key, (sad[i] << 16) | (uint32_t)(D - 1 - (d0 + 2 * i)));
          key = min(key, (sad[i] & 0xFFFF0000u) | (uint32_t)(D - 1 - (d0 + 2 * i + 1)));
        }
      }
      key = __reduce_min_sync(gmask, key);
      const int minsad = (int)(key >> 16), mind = (int)(key & 0xFFFFu);
      // neighbours in k order with the mirrored ends sad[-1]=sad[1], sad[D]=sad[D-2]
      const int kp = mind + 1 < D ? mind + 1 : D - 2, kn = mind - 1 >= 0 ? mind - 1 : 1;
      const int dp = D - 1 - kp, dn = D - 1 - kn;
      uint32_t vp = 0, vn = 0;
      bool rej = false;
      if (on) {
        const int thresh = minsad + (minsad * cfg.uniq / 100);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int d = d0 + i, k = D - 1 - d;
          const int s = (int)((i & 1) ? (sad[i >> 1] >> 16) : (sad[i >> 1] & 0xFFFFu));
          if (d == dp) vp = s;
          if (d == dn) vn = s;
          if (cfg.uniq > 0 && (k < mind - 1 || k > mind + 1) && s <= thresh) rej = true;
        }
      }
      vp = __reduce_max_sync(gmask, vp);
      vn = __reduce_max_sync(gmask, vn);
      const bool anyRej = (__ballot_sync(0xFFFFFFFFu, rej) & gmask) != 0;
      if (on && gl == 0) {
        int out = cfg.filtered;
        if (tex >= cfg.tex && !anyRej) {
          const int p = (int)vp, n = (int)vn;
          const int dd = p + n - 2 * minsad + abs(p - n);
          out = ((D - mind - 1 + cfg.minD) * 256 + (dd != 0 ? (p - n) * 256 / dd : 0) + 15) >> 4;
        }
        *(int16_t*)((uint8_t*)disp + (size_t)y * dstride + (size_t)x * 2) = (int16_t)out;
      }
    }
  }
}

__global__ void k_fill_rows(int16_t* p, size_t stride, int W, int H, int v) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < W) *(int16_t*)((uint8_t*)p + (size_t)y * stride + (size_t)x * 2) = (int16_t)v;
}

size_t bm_smem_bytes(int D, int blockSize) { return bm_smem_layout(D, blockSize / 2).total; }

cudaError_t bm_forward(const BmCfg& cfg, const uint8_t* left, size_t lstride, const uint8_t* right, size_t rstride,
                       uint8_t* pfL, uint8_t* pfR, int16_t* disp, size_t dstride, cudaStream_t st, int* nlaunch) {
  const int W = cfg.W, H = cfg.H, r = cfg.bs / 2;
  {
    dim3 g((W + 255) / 256, H);
    k_fill_rows<<<g, 256, 0, st>>>(disp, dstride, W, H, cfg.filtered);
    (*nlaunch)++;
  }
  const int W1 = W - std::max(cfg.D - 1 + cfg.minD, 0);
  if (W1 - 2 * r + std::min(cfg.minD, 0) <= 0 || H - 2 * r <= 0) return cudaGetLastError();
  {
    dim3 g((W + 255) / 256, H, 2);
    k_bm_prefilter<<<g, 256, 0, st>>>(left, lstride, right, rstride, pfL, pfR, W, H, cfg.cap);
    (*nlaunch)++;
  }
  const BmSmem L = bm_smem_layout(cfg.D, r);
  const int tiles = (W1 + kBmTW - 1) / kBmTW;
  const int outRows = H - 2 * r;
  int bands = std::max(1, std::min((148 * 4 + tiles - 1) / tiles, (outRows + 63) / 64));
  const int rowsPerBand = (outRows + bands - 1) / bands;
  bands = (outRows + rowsPerBand - 1) / rowsPerBand;
  dim3 g(tiles, bands);
  cudaError_t err = cudaSuccess;
  auto go = [&](auto lpp) {
    constexpr int LPP = decltype(lpp)::value;
    err = cudaFuncSetAttribute(k_bm_sad<LPP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    if (err != cudaSuccess) return;
    k_bm_sad<LPP><<<g, kBmThreads, L.total, st>>>(pfL, pfR, disp, dstride, cfg, rowsPerBand);
    err = cudaGetLastError();
  };
  const int v = cfg.D / 8;
  if (v <= 2) go(std::integral_constant<int, 2>());
  else if (v <= 4) go(std::integral_constant<int, 4>());
  else if (v <= 8) go(std::integral_constant<int, 8>());
  else if (v <= 16) go(std::integral_constant<int, 16>());
  else go(std::integral_constant<int, 32>());
  (*nlaunch)++;
  return err;
}

}  // namespace svgpu
