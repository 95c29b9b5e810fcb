// table_lock.h -- a __constant__ table is per device and shared by every stream and context of
// the process, so its users are serialised: the upload of the next job waits (on the device, via
// an event) for the kernel of the previous one, and the host-side sequence upload -> launch ->
// record is atomic under a mutex.  The reference has the same hazard (file-scope __constant__
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
    cudaEvent_t last_use[kMaxDevices] = {};
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
        // always ordered after the table's previous user: its upload may still be in flight on
        // another stream (free when it is the same stream)
        if (lock_.used[device_])
            status_ = cudaStreamWaitEvent(stream_, lock_.last_use[device_], 0);
    }
    bool needs_upload() const { return upload_; }
    void uploaded() { if (status_ == cudaSuccess) lock_.resident[device_].assign(bytes_, bytes_ + size_); }
    // the upload or launch failed: the device copy is unknown
    void invalidate() { if (device_ >= 0 && device_ < TableLock::kMaxDevices) lock_.resident[device_].clear(); }
    ~TableUse()
    {
        if (status_ == cudaSuccess) {
            if (!lock_.used[device_]) {
                if (cudaEventCreateWithFlags(&lock_.last_use[device_], cudaEventDisableTiming) == cudaSuccess)
                    lock_.used[device_] = true;
            }
            if (lock_.used[device_])
                cudaEventRecord(lock_.last_use[device_], stream_);
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
