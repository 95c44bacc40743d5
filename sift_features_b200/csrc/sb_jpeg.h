// sb_jpeg.h -- nvJPEG front end of the JPEG entry points (host code only).
//
// The reference's callers decode a JPEG and convert it to gray before sift() (examples/run-sift.rs:8,
// examples/sift-match.rs:49, src/lib.rs:1012).  Here the bitstreams of a group are decoded on the device with
// nvJPEG's batched API straight into the slot's input buffers, so the host never touches pixels.  nvJPEG is library
// code and is loaded with dlopen on first use: libsift_b200.so has no link-time dependency on it, and contexts that
// never see a JPEG never load it.
//
// Backends are tried in this order: the hardware JPEG engines, GPU-assisted Huffman decoding, nvJPEG's default
// (Huffman decoding on the calling thread).  A backend that cannot be created on this device or refuses a batch
// (progressive streams on the hardware engines, for instance) passes the batch on to the next one;
// SB200_JPEG_BACKEND=hardware|gpu|default pins one.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvjpeg.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace sb {

struct NvJpegApi {
    void* lib = nullptr;
    decltype(&nvjpegCreateEx) CreateEx = nullptr;
    decltype(&nvjpegDestroy) Destroy = nullptr;
    decltype(&nvjpegJpegStateCreate) JpegStateCreate = nullptr;
    decltype(&nvjpegJpegStateDestroy) JpegStateDestroy = nullptr;
    decltype(&nvjpegGetImageInfo) GetImageInfo = nullptr;
    decltype(&nvjpegDecodeBatchedInitialize) DecodeBatchedInitialize = nullptr;
    decltype(&nvjpegDecodeBatched) DecodeBatched = nullptr;

    bool load(std::string& err) {
        if (lib) return true;
        const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"};
        for (const char* n : names)
            if ((lib = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
        if (!lib) {
            err = std::string("nvJPEG is not available: ") + dlerror();
            return false;
        }
        bool ok = true;
        auto sym = [&](const char* name) {
            void* p = dlsym(lib, name);
            if (!p) { ok = false; err = std::string("nvJPEG lacks ") + name; }
            return p;
        };
        CreateEx = reinterpret_cast<decltype(CreateEx)>(sym("nvjpegCreateEx"));
        Destroy = reinterpret_cast<decltype(Destroy)>(sym("nvjpegDestroy"));
        JpegStateCreate = reinterpret_cast<decltype(JpegStateCreate)>(sym("nvjpegJpegStateCreate"));
        JpegStateDestroy = reinterpret_cast<decltype(JpegStateDestroy)>(sym("nvjpegJpegStateDestroy"));
        GetImageInfo = reinterpret_cast<decltype(GetImageInfo)>(sym("nvjpegGetImageInfo"));
        DecodeBatchedInitialize = reinterpret_cast<decltype(DecodeBatchedInitialize)>(sym("nvjpegDecodeBatchedInitialize"));
        DecodeBatched = reinterpret_cast<decltype(DecodeBatched)>(sym("nvjpegDecodeBatched"));
        if (!ok) { dlclose(lib); lib = nullptr; }
        return ok;
    }
};

inline const char* nvjpeg_status_name(nvjpegStatus_t s) {
    switch (s) {
        case NVJPEG_STATUS_SUCCESS: return "success";
        case NVJPEG_STATUS_NOT_INITIALIZED: return "not initialized";
        case NVJPEG_STATUS_INVALID_PARAMETER: return "invalid parameter";
        case NVJPEG_STATUS_BAD_JPEG: return "bad jpeg";
        case NVJPEG_STATUS_JPEG_NOT_SUPPORTED: return "jpeg not supported";
        case NVJPEG_STATUS_ALLOCATOR_FAILURE: return "allocator failure";
        case NVJPEG_STATUS_EXECUTION_FAILED: return "execution failed";
        case NVJPEG_STATUS_ARCH_MISMATCH: return "arch mismatch";
        case NVJPEG_STATUS_INTERNAL_ERROR: return "internal error";
        case NVJPEG_STATUS_IMPLEMENTATION_NOT_SUPPORTED: return "implementation not supported";
        case NVJPEG_STATUS_INCOMPLETE_BITSTREAM: return "incomplete bitstream";
    }
    return "?";
}

// one decoder per context; `lanes` independent decode states (one per slot) so that groups in flight do not share
// nvJPEG's staging buffers
class JpegDecoder {
  public:
    struct Info { uint32_t w = 0, h = 0, components = 0; };

    ~JpegDecoder() { release(); }

    bool init(int lanes, std::string& err) {
        if (!engines_.empty()) return true;
        if (!api_.load(err)) return false;
        lanes_ = lanes;
        const char* pin = getenv("SB200_JPEG_BACKEND");
        auto want = [&](const char* name) { return !pin || !*pin || !strcmp(pin, name); };
        if (want("hardware")) engines_.push_back({NVJPEG_BACKEND_HARDWARE, "hardware"});
        if (want("gpu")) engines_.push_back({NVJPEG_BACKEND_GPU_HYBRID, "gpu"});
        if (want("default")) engines_.push_back({NVJPEG_BACKEND_DEFAULT, "default"});
        if (engines_.empty()) { err = "SB200_JPEG_BACKEND must be hardware, gpu or default"; return false; }
        return true;
    }

    // header only; no device work
    bool info(const uint8_t* data, size_t len, Info& out, std::string& err) {
        Engine* e = any_engine(err);
        if (!e) return false;
        int comps = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
        nvjpegChromaSubsampling_t sub;
        nvjpegStatus_t st = api_.GetImageInfo(e->handle, data, len, &comps, &sub, ws, hs);
        if (st != NVJPEG_STATUS_SUCCESS) { err = std::string("not a decodable JPEG: ") + nvjpeg_status_name(st); return false; }
        if (ws[0] <= 0 || hs[0] <= 0 || (comps != 1 && comps != 3)) { err = "unsupported JPEG (need 1 or 3 components)"; return false; }
        out.w = (uint32_t)ws[0];
        out.h = (uint32_t)hs[0];
        out.components = (uint32_t)comps;
        return true;
    }

    // decodes n bitstreams into dst[i] (pitch bytes per row): interleaved RGB when rgb, else the Y plane
    bool decode(int lane, const uint8_t* const* data, const size_t* lens, uint32_t n, bool rgb, uint8_t* const* dst, size_t pitch,
                cudaStream_t stream, std::string& err) {
        const nvjpegOutputFormat_t fmt = rgb ? NVJPEG_OUTPUT_RGBI : NVJPEG_OUTPUT_Y;
        std::vector<nvjpegImage_t> outs(n);
        for (uint32_t i = 0; i < n; i++) {
            memset(&outs[i], 0, sizeof(nvjpegImage_t));
            outs[i].channel[0] = dst[i];
            outs[i].pitch[0] = pitch;
        }
        std::string why;
        for (size_t k = first_; k < engines_.size(); k++) {
            Engine& e = engines_[k];
            if (!create(e)) { why += std::string(" [") + e.name + ": unavailable]"; continue; }
            Lane& l = e.lane[lane];
            nvjpegStatus_t st = NVJPEG_STATUS_SUCCESS;
            if (l.batch != (int)n || l.fmt != fmt) {
                st = api_.DecodeBatchedInitialize(e.handle, l.state, (int)n, 1, fmt);
                l.batch = st == NVJPEG_STATUS_SUCCESS ? (int)n : -1;
                l.fmt = fmt;
            }
            if (st == NVJPEG_STATUS_SUCCESS) {
                st = api_.DecodeBatched(e.handle, l.state, data, lens, outs.data(), stream);
                if (st == NVJPEG_STATUS_SUCCESS) { used_ = e.name; return true; }
                l.batch = -1;   // a failed batch has to be initialised again
            }
            cudaGetLastError();
            why += std::string(" [") + e.name + ": " + nvjpeg_status_name(st) + "]";
            if (st == NVJPEG_STATUS_BAD_JPEG || st == NVJPEG_STATUS_INCOMPLETE_BITSTREAM) break;   // the data, not the backend
        }
        err = "JPEG decode failed:" + why;
        return false;
    }

    const char* backend() const { return used_; }

    void release() {
        for (auto& e : engines_) {
            for (auto& l : e.lane) if (l.state) api_.JpegStateDestroy(l.state);
            if (e.handle) api_.Destroy(e.handle);
            e.lane.clear();
            e.handle = nullptr;
        }
        engines_.clear();
    }

  private:
    struct Lane { nvjpegJpegState_t state = nullptr; int batch = -1; nvjpegOutputFormat_t fmt = NVJPEG_OUTPUT_Y; };
    struct Engine {
        nvjpegBackend_t backend;
        const char* name;
        nvjpegHandle_t handle = nullptr;
        bool dead = false;
        std::vector<Lane> lane;
    };

    bool create(Engine& e) {
        if (e.handle) return true;
        if (e.dead) return false;
        if (api_.CreateEx(e.backend, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &e.handle) != NVJPEG_STATUS_SUCCESS) {
            e.handle = nullptr; e.dead = true; cudaGetLastError();
            return false;
        }
        e.lane.resize(lanes_);
        for (auto& l : e.lane)
            if (api_.JpegStateCreate(e.handle, &l.state) != NVJPEG_STATUS_SUCCESS) {
                for (auto& m : e.lane) if (m.state) api_.JpegStateDestroy(m.state);
                e.lane.clear();
                api_.Destroy(e.handle);
                e.handle = nullptr; e.dead = true; cudaGetLastError();
                return false;
            }
        return true;
    }

    Engine* any_engine(std::string& err) {
        for (auto& e : engines_) if (create(e)) return &e;
        err = "no nvJPEG backend could be created on this device";
        return nullptr;
    }

    NvJpegApi api_;
    std::vector<Engine> engines_;
    size_t first_ = 0;
    int lanes_ = 1;
    const char* used_ = "none";
};

}  // namespace sb
