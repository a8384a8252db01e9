// Host-side plumbing shared by the C-ABI wrappers: error reporting and staging of host buffers.
#pragma once

#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/pttools_b200.h"

namespace ptt {

char* error_buffer();   // defined in ptt_api.cu (thread-local)

inline int set_error(const int code, const char* msg) {
    snprintf(error_buffer(), 512, "%s", msg);
    return code;
}

inline int set_cuda_error(const char* where, const cudaError_t e) {
    snprintf(error_buffer(), 512, "%s: %s", where, cudaGetErrorString(e));
    return PTT_E_CUDA;
}

// With PTT_MEM_DEVICE the caller's pointers are used as they are and nothing is synchronised.
// With PTT_MEM_HOST every input is copied to a temporary device buffer and every output is copied back in
// finish(), which also synchronises the stream.
class Stager {
public:
    Stager(const int mem, cudaStream_t stream) : host_(mem == PTT_MEM_HOST), stream_(stream), err_(cudaSuccess) {}
    ~Stager() {
        for (void* p : bufs_) cudaFree(p);
    }
    template <class T>
    const T* in(const T* p, const size_t n) {
        if (!host_) return p;
        T* d = alloc<T>(n);
        if (d) check(cudaMemcpyAsync(d, p, n * sizeof(T), cudaMemcpyHostToDevice, stream_));
        return d;
    }
    template <class T>
    T* out(T* p, const size_t n) {
        if (!host_) return p;
        T* d = alloc<T>(n);
        if (d) backs_.push_back({p, d, n * sizeof(T)});
        return d;
    }
    template <class T>
    T* inout(T* p, const size_t n) {
        if (!host_) return p;
        T* d = alloc<T>(n);
        if (d) {
            check(cudaMemcpyAsync(d, p, n * sizeof(T), cudaMemcpyHostToDevice, stream_));
            backs_.push_back({p, d, n * sizeof(T)});
        }
        return d;
    }
    template <class T>
    T* scratch(const size_t n) { return alloc<T>(n); }
    bool ok() const { return err_ == cudaSuccess; }
    int fail() const { return set_cuda_error("staging", err_); }
    int finish(const char* what) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return set_cuda_error(what, e);
        if (host_) {
            for (const Back& b : backs_) {
                e = cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, stream_);
                if (e != cudaSuccess) return set_cuda_error(what, e);
            }
            e = cudaStreamSynchronize(stream_);
            if (e != cudaSuccess) return set_cuda_error(what, e);
        }
        return PTT_OK;
    }

private:
    struct Back {
        void* host;
        void* dev;
        size_t bytes;
    };
    template <class T>
    T* alloc(const size_t n) {
        void* d = nullptr;
        const cudaError_t e = cudaMalloc(&d, (n ? n : 1) * sizeof(T));
        if (e != cudaSuccess) {
            err_ = e;
            return nullptr;
        }
        bufs_.push_back(d);
        return (T*)d;
    }
    void check(const cudaError_t e) {
        if (e != cudaSuccess && err_ == cudaSuccess) err_ = e;
    }
    bool host_;
    cudaStream_t stream_;
    cudaError_t err_;
    std::vector<void*> bufs_;
    std::vector<Back> backs_;
};

// ptt_a2_batch with an explicit row stride of the A^2 table (ptt_ssm.cu); used by the sweep driver
int a2_batch_strided(const ptt_a2_plan_t* plan, int P, const double* rows_v, const double* rows_w, const double* rows_xi,
                     int64_t row_stride, const int32_t* off, const int32_t* npts, const double* pp, double* work,
                     int64_t work_stride, double* a2, int64_t a2_stride, double* fp2_2, double* lam2, void* stream);

}  // namespace ptt
