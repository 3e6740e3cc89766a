// shim/cuda_executor.cpp — gko::CudaExecutor members for the B200 drop-in libginkgo_cuda.so.
// Our implementation of the non-inline members the reference defines in
// cuda/base/executor.cpp:59-293 and cuda/base/scoped_device_id.cpp (written against the
// CUDA runtime only: this build has no cuBLAS / cuSPARSE handles because no kernel here
// uses a vendor library).
#include <cuda_runtime.h>

#include <iostream>
#include <memory>
#include <string>

#include <ginkgo/core/base/device.hpp>
#include <ginkgo/core/base/exception_helpers.hpp>
#include <ginkgo/core/base/executor.hpp>
#include <ginkgo/core/base/scoped_device_id_guard.hpp>
#include <ginkgo/core/base/version.hpp>

namespace gko {
namespace {

#define SHIM_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) throw CudaError(__FILE__, __LINE__, #call, e__); \
    } while (0)

// RAII device switch (role of detail::cuda_scoped_device_id_guard)
class device_guard : public detail::generic_scoped_device_id_guard {
public:
    explicit device_guard(int id)
    {
        SHIM_CUDA(cudaGetDevice(&orig_));
        if (orig_ != id) {
            SHIM_CUDA(cudaSetDevice(id));
            reset_ = true;
        }
    }
    ~device_guard() override
    {
        if (reset_ && cudaSetDevice(orig_) != cudaSuccess) {
            std::cerr << "gko_b200 shim: cannot restore CUDA device " << orig_ << std::endl;
            std::exit(1);
        }
    }

private:
    int orig_ = 0;
    bool reset_ = false;
};

}  // namespace

version version_info::get_cuda_version() noexcept { return {GKO_VERSION_STR, "b200-native"}; }

std::shared_ptr<CudaExecutor> CudaExecutor::create(int device_id, std::shared_ptr<Executor> master, bool device_reset,
                                                   allocation_mode alloc_mode)
{
    return std::shared_ptr<CudaExecutor>(new CudaExecutor(device_id, std::move(master), device_reset, alloc_mode));
}

void CudaExecutor::populate_exec_info(const machine_topology*) {}

void OmpExecutor::raw_copy_to(const CudaExecutor* dest, size_type num_bytes, const void* src, void* dst) const
{
    if (num_bytes > 0) {
        device_guard g(dest->get_device_id());
        SHIM_CUDA(cudaMemcpy(dst, src, num_bytes, cudaMemcpyHostToDevice));
    }
}

void CudaExecutor::raw_free(void* ptr) const noexcept
{
    device_guard g(this->get_device_id());
    if (cudaFree(ptr) != cudaSuccess) {
        std::cerr << "gko_b200 shim: unrecoverable CUDA error in cudaFree" << std::endl;
        std::exit(1);  // raw_free must not throw (cuda/base/executor.cpp:112-128)
    }
}

void* CudaExecutor::raw_alloc(size_type num_bytes) const
{
    void* p = nullptr;
    device_guard g(this->get_device_id());
    const cudaError_t e = cudaMalloc(&p, num_bytes);
    if (e == cudaErrorMemoryAllocation) throw AllocationError(__FILE__, __LINE__, "cuda", num_bytes);
    SHIM_CUDA(e);
    return p;
}

void CudaExecutor::raw_copy_to(const OmpExecutor*, size_type num_bytes, const void* src, void* dst) const
{
    if (num_bytes > 0) {
        device_guard g(this->get_device_id());
        SHIM_CUDA(cudaMemcpy(dst, src, num_bytes, cudaMemcpyDeviceToHost));
    }
}

void CudaExecutor::raw_copy_to(const CudaExecutor* dest, size_type num_bytes, const void* src, void* dst) const
{
    if (num_bytes > 0) {
        device_guard g(this->get_device_id());
        SHIM_CUDA(cudaMemcpyPeer(dst, dest->get_device_id(), src, this->get_device_id(), num_bytes));
    }
}

void CudaExecutor::raw_copy_to(const HipExecutor*, size_type, const void*, void*) const GKO_NOT_SUPPORTED(this);
void CudaExecutor::raw_copy_to(const DpcppExecutor*, size_type, const void*, void*) const GKO_NOT_SUPPORTED(this);

void CudaExecutor::synchronize() const
{
    device_guard g(this->get_device_id());
    SHIM_CUDA(cudaDeviceSynchronize());
}

scoped_device_id_guard CudaExecutor::get_scoped_device_id_guard() const { return {this, this->get_device_id()}; }

void CudaExecutor::run(const Operation& op) const
{
    this->template log<log::Logger::operation_launched>(this, &op);
    device_guard g(this->get_device_id());
    op.run(std::static_pointer_cast<const CudaExecutor>(this->shared_from_this()));
    this->template log<log::Logger::operation_completed>(this, &op);
}

std::string CudaError::get_error(int64 code)
{
    return std::string(cudaGetErrorName(static_cast<cudaError_t>(code))) + ": " +
           cudaGetErrorString(static_cast<cudaError_t>(code));
}

int CudaExecutor::get_num_devices()
{
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if (e == cudaErrorNoDevice) return 0;
    SHIM_CUDA(e);
    return n;
}

void CudaExecutor::set_gpu_property()
{
    const int id = this->get_device_id();
    if (id < 0 || id >= get_num_devices()) return;
    device_guard g(id);
    auto& info = this->get_exec_info();
    SHIM_CUDA(cudaDeviceGetAttribute(&info.major, cudaDevAttrComputeCapabilityMajor, id));
    SHIM_CUDA(cudaDeviceGetAttribute(&info.minor, cudaDevAttrComputeCapabilityMinor, id));
    SHIM_CUDA(cudaDeviceGetAttribute(&info.num_computing_units, cudaDevAttrMultiProcessorCount, id));
    int max_threads = 0;
    SHIM_CUDA(cudaDeviceGetAttribute(&max_threads, cudaDevAttrMaxThreadsPerBlock, id));
    info.max_workgroup_size = max_threads;
    info.max_workitem_sizes = {1024, 1024, 64};
    info.num_pu_per_cu = 4;
    info.max_subgroup_size = 32;
}

void CudaExecutor::init_handles() {}  // no cuBLAS / cuSPARSE on this path

scoped_device_id_guard::scoped_device_id_guard(const CudaExecutor*, int device_id)
    : scope_(std::make_unique<device_guard>(device_id))
{}

}  // namespace gko
