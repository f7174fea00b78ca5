// Input pipeline, device side: JPEG bitstreams -> planar RGB uint8 (n, 3, H, W) in HBM, i.e. exactly what the
// engine's patch-embed / stem kernels read.  Replaces the host decode + ToTensorV2 of the reference's dataset classes
// (support.py:73-81, datasets/road_anomaly.py, datasets/fishyscapes.py:19-63: PIL / cv2 on the main thread) for JPEG inputs.
// The decoder is nvJPEG (a CUDA toolkit library, like cuBLAS: not a kernel of this repo); it is bound with dlopen so that the
// library loads -- and every other entry point works -- on a box without libnvjpeg.  PNG inputs stay on the host path
// (inflate is serial per stream): rba_b200.pipeline.PinnedBatcher.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <map>
#include <mutex>

#include "common.cuh"

namespace rba {

struct NvJpegApi {
  void* so = nullptr;
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*,
                           cudaStream_t) = nullptr;
  bool ok = false;
};

static NvJpegApi& nvjpeg_api() {
  static NvJpegApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so"}) {
      api.so = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (api.so) break;
    }
    if (!api.so) return;
    api.CreateSimple = (decltype(api.CreateSimple))dlsym(api.so, "nvjpegCreateSimple");
    api.JpegStateCreate = (decltype(api.JpegStateCreate))dlsym(api.so, "nvjpegJpegStateCreate");
    api.GetImageInfo = (decltype(api.GetImageInfo))dlsym(api.so, "nvjpegGetImageInfo");
    api.Decode = (decltype(api.Decode))dlsym(api.so, "nvjpegDecode");
    api.ok = api.CreateSimple && api.JpegStateCreate && api.GetImageInfo && api.Decode;
  });
  return api;
}

struct NvJpegCtx {
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
};
static std::mutex g_jpeg_mu;                       // nvjpegDecode with one state is not re-entrant
static std::map<int, NvJpegCtx> g_jpeg_ctx;        // per device

static int jpeg_ctx(NvJpegCtx** out) {
  NvJpegApi& api = nvjpeg_api();
  if (!api.ok) return fail(RBA_ERR_CUDA, "nvJPEG is not available (libnvjpeg.so.12 could not be loaded)");
  int dev = 0;
  RBA_CUDA(cudaGetDevice(&dev));
  NvJpegCtx& c = g_jpeg_ctx[dev];
  if (!c.handle) {
    if (api.CreateSimple(&c.handle) != NVJPEG_STATUS_SUCCESS) return fail(RBA_ERR_CUDA, "nvjpegCreateSimple failed");
    if (api.JpegStateCreate(c.handle, &c.state) != NVJPEG_STATUS_SUCCESS) return fail(RBA_ERR_CUDA, "nvjpegJpegStateCreate failed");
  }
  *out = &c;
  return RBA_OK;
}

}  // namespace rba

extern "C" int rba_jpeg_available(void) { return rba::nvjpeg_api().ok ? 1 : 0; }

extern "C" int rba_jpeg_info(const uint8_t* data, int64_t nbytes, int* height, int* width, int* channels) {
  using namespace rba;
  RBA_CHECK(data && nbytes > 0 && height && width && channels, "rba_jpeg_info: null pointer / empty stream");
  std::lock_guard<std::mutex> lk(g_jpeg_mu);
  NvJpegCtx* c = nullptr;
  RBA_TRY_(jpeg_ctx(&c));
  int nc = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
  nvjpegChromaSubsampling_t ss;
  const nvjpegStatus_t st = nvjpeg_api().GetImageInfo(c->handle, data, (size_t)nbytes, &nc, &ss, ws, hs);
  RBA_CHECK(st == NVJPEG_STATUS_SUCCESS, "rba_jpeg_info: not a decodable JPEG stream (nvjpeg status %d)", (int)st);
  *height = hs[0]; *width = ws[0]; *channels = nc;
  return RBA_OK;
}

// data[i] / nbytes[i]: n JPEG bitstreams in HOST memory, all H x W; out: device (n, 3, H, W) uint8, planar RGB (grey streams are
// expanded).  Stream-ordered: the planes are valid for work submitted to `stream` after the call.
extern "C" int rba_jpeg_decode(const uint8_t* const* data, const int64_t* nbytes, int n, uint8_t* out, int H, int W, void* stream) {
  using namespace rba;
  if (n == 0) return RBA_OK;
  RBA_CHECK(data && nbytes && out && n > 0 && H > 0 && W > 0, "rba_jpeg_decode: bad arguments");
  std::lock_guard<std::mutex> lk(g_jpeg_mu);
  NvJpegCtx* c = nullptr;
  RBA_TRY_(jpeg_ctx(&c));
  NvJpegApi& api = nvjpeg_api();
  for (int i = 0; i < n; ++i) {
    RBA_CHECK(data[i] && nbytes[i] > 0, "rba_jpeg_decode: stream %d is empty", i);
    int nc = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t ss;
    nvjpegStatus_t st = api.GetImageInfo(c->handle, data[i], (size_t)nbytes[i], &nc, &ss, ws, hs);
    RBA_CHECK(st == NVJPEG_STATUS_SUCCESS, "rba_jpeg_decode: stream %d is not a decodable JPEG (nvjpeg status %d)", i, (int)st);
    RBA_CHECK(hs[0] == H && ws[0] == W, "rba_jpeg_decode: stream %d is %d x %d, expected %d x %d", i, hs[0], ws[0], H, W);
    nvjpegImage_t img;
    memset(&img, 0, sizeof(img));
    for (int ch = 0; ch < 3; ++ch) {
      img.channel[ch] = out + ((size_t)i * 3 + ch) * H * W;
      img.pitch[ch] = (size_t)W;
    }
    st = api.Decode(c->handle, c->state, data[i], (size_t)nbytes[i], NVJPEG_OUTPUT_RGB, &img, (cudaStream_t)stream);
    RBA_CHECK(st == NVJPEG_STATUS_SUCCESS, "rba_jpeg_decode: nvjpegDecode failed on stream %d (status %d)", i, (int)st);
  }
  return RBA_OK;
}
