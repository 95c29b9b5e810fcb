// table_lock.h -- a __constant__ table is per device and shared by every stream and context of
// the process, so its users are serialised: the host-side sequence upload -> launch is atomic under a mutex, and a
// user on ANOTHER stream than the previous one first waits (on the host) for that stream, whose kernel may still be
// reading the table.  Users that follow each other on one stream -- the common case -- are ordered by the stream itself
// and pay nothing (the first version recorded and waited on an event around every launch: two more stream
// operations on a 15 us call).  The reference has the same hazard (file-scope __constant__
// OPTION / MOPTION, DP/MonteCarloKernel.cu:53-59) and no protection.
#pragma once

#include <cuda_runtime.h>
#include <cstring>
#include <mutex>
#include <vector>

namespace mcb {

struct TableLock {
    static constexpr int kMaxDevices = 64;
    std::mutex mu;
    cudaStream_t last_stream[kMaxDevices] = {};
    bool used[kMaxDevices] = {};
    // what the device's table holds once everything enqueued so far has run: a job repeated with the
    // same parameters skips the upload
    std::vector<unsigned char> resident[kMaxDevices];
};

class TableUse {
public:
    // `bytes`/`size`: the table image this job needs.  needs_upload() tells the caller whether the device copy
    // differs; after the upload has been ENQUEUED successfully the caller says so with uploaded() -- only then does the
    // cache claim the new image (a failed wait or copy must not leave it describing a table the device never got).
    TableUse(TableLock &lock, cudaStream_t stream, const void *bytes, size_t size)
        : lock_(lock), stream_(stream), bytes_((const unsigned char *)bytes), size_(size)
    {
        lock_.mu.lock();
        status_ = cudaGetDevice(&device_);
        if (status_ == cudaSuccess && (device_ < 0 || device_ >= TableLock::kMaxDevices))
            status_ = cudaErrorInvalidDevice;
        if (status_ != cudaSuccess)
            return;
        const std::vector<unsigned char> &res = lock_.resident[device_];
        upload_ = res.size() != size || std::memcmp(res.data(), bytes, size) != 0;
        // the table's previous user ran on another stream: let it finish (if that stream is gone, so is its work)
        if (lock_.used[device_] && lock_.last_stream[device_] != stream_ && cudaStreamSynchronize(lock_.last_stream[device_]) != cudaSuccess)
            cudaGetLastError();
    }
    bool needs_upload() const { return upload_; }
    void uploaded() { if (status_ == cudaSuccess) lock_.resident[device_].assign(bytes_, bytes_ + size_); }
    // the upload or launch failed: the device copy is unknown
    void invalidate() { if (device_ >= 0 && device_ < TableLock::kMaxDevices) lock_.resident[device_].clear(); }
    ~TableUse()
    {
        if (status_ == cudaSuccess) {
            lock_.last_stream[device_] = stream_;
            lock_.used[device_] = true;
        }
        lock_.mu.unlock();
    }
    cudaError_t status() const { return status_; }

private:
    TableLock &lock_;
    cudaStream_t stream_;
    const unsigned char *bytes_;
    size_t size_;
    int device_ = -1;
    cudaError_t status_ = cudaSuccess;
    bool upload_ = true;
};

}  // namespace mcb
