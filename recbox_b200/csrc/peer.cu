// Peer-visible device memory for the row-sharded table (include/recbox_b200.h "(e)"): plain
// cudaMalloc blocks exported / mapped with CUDA IPC so that each rank's fused kernels can address
// every other rank's table and gradient shard over NVLink / NVSwitch.
#include <string.h>
#include "rbx_common.cuh"

static_assert(sizeof(cudaIpcMemHandle_t) == RBX_PEER_HANDLE_BYTES, "handle size");

#define RBX_CUDA(call, who)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            cudaGetLastError();                                                               \
            return rbx_fail(RBX_ERR_CUDA, "%s: %s failed: %s", who, #call, cudaGetErrorString(e__)); \
        }                                                                                     \
    } while (0)

extern "C" {

int rbx_peer_alloc(size_t bytes, void** ptr) {
    const char* who = "rbx_peer_alloc";
    RBX_RANGE(who);
    RBX_REQUIRE(ptr != nullptr, "%s: null out pointer", who);
    *ptr = nullptr;
    if (bytes == 0) bytes = 16;
    RBX_CUDA(cudaMalloc(ptr, bytes), who);
    return RBX_OK;
}

int rbx_peer_free(void* ptr) {
    if (ptr) RBX_CUDA(cudaFree(ptr), "rbx_peer_free");
    return RBX_OK;
}

int rbx_peer_export(const void* ptr, unsigned char handle[RBX_PEER_HANDLE_BYTES]) {
    const char* who = "rbx_peer_export";
    RBX_RANGE(who);
    RBX_REQUIRE(ptr && handle, "%s: null pointer", who);
    cudaIpcMemHandle_t h;
    RBX_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)), who);
    memcpy(handle, &h, sizeof(h));
    return RBX_OK;
}

int rbx_peer_open(const unsigned char handle[RBX_PEER_HANDLE_BYTES], void** ptr) {
    const char* who = "rbx_peer_open";
    RBX_RANGE(who);
    RBX_REQUIRE(ptr && handle, "%s: null pointer", who);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    *ptr = nullptr;
    RBX_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess), who);
    return RBX_OK;
}

int rbx_peer_close(void* ptr) {
    if (ptr) RBX_CUDA(cudaIpcCloseMemHandle(ptr), "rbx_peer_close");
    return RBX_OK;
}

int rbx_peer_can_access(int device, int peer_device) {
    int ok = 0;
    RBX_CUDA(cudaDeviceCanAccessPeer(&ok, device, peer_device), "rbx_peer_can_access");
    return ok;
}

}  // extern "C"
