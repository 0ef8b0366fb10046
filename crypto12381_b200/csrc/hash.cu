// Batched SHA3-512 and hash-to-Zp: __global__ wrappers + C-ABI entries (include/c12381_cuda.h); bodies in hash.cuh.
#include "msm_impl.cuh"
#include "hash.cuh"

namespace c12 {

// mode 0: 64-byte digests; mode 1: digests reduced mod r, 32 bytes big-endian
__global__ void __launch_bounds__(128) k_sha3_512(const uint8_t* __restrict__ msgs, size_t len, uint32_t B, int mode, uint8_t* __restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    if (mode == 0)
        sha3_512(msgs + (size_t)i * len, len, out + 64ull * i);
    else
        hash_to_zp_body(msgs + (size_t)i * len, len, out + 32ull * i);
}

// mode 2: hash each message to G1 (49 bytes compressed); mode 3: the messages are field elements (48 B big-endian), mapped to G1
__global__ void __launch_bounds__(128) k_hash_to_g1(const uint8_t* __restrict__ msgs, size_t len, uint32_t B, int mode, uint8_t* __restrict__ out,
                                                    int* flags)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    if (mode == 2)
        hash_to_g1_body(msgs + (size_t)i * len, len, out + 49ull * i);
    else if (!map_to_g1_body(msgs + 48ull * i, out + 49ull * i))
        atomicOr(flags, FLAG_BAD_POINT);
}

static int hash_run(const uint8_t* d_msgs, size_t len, size_t B, int mode, uint8_t* d_out, cudaStream_t s)
{
    if (B == 0) return C12381_OK;
    if (B > 0x7fffffffull) return set_error(C12381_EARG, "hash: too many messages");
    if (mode < 2)
        k_sha3_512<<<cdiv(B, 128), 128, 0, s>>>(d_msgs, len, (uint32_t)B, mode, d_out);
    else
        k_hash_to_g1<<<cdiv(B, 128), 128, 0, s>>>(d_msgs, len, (uint32_t)B, mode, d_out, flags_word());
    C12_LAUNCHED();
    return C12381_OK;
}

static size_t out_size(int mode) { return mode == 0 ? 64 : mode == 1 ? 32 : 49; }

static int hash_host(const uint8_t* msgs, size_t len, size_t B, int mode, uint8_t* out)
{
    C12_REQUIRE_CTX();
    if (B && (!out || (len && !msgs))) return set_error(C12381_EARG, "hash: null pointer");
    const void* in[1] = {msgs};
    size_t sz[1] = {B * len};
    return with_staged(in, sz, 1, out, B * out_size(mode), 0, [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) {
        return hash_run(d_in[0], len, B, mode, d_out, s);
    });
}

static int hash_dev(const uint8_t* msgs, size_t len, size_t B, int mode, uint8_t* out, void* stream)
{
    C12_REQUIRE_CTX();
    if (B && (!out || (len && !msgs))) return set_error(C12381_EARG, "hash: null pointer");
    return hash_run(msgs, len, B, mode, out, pick_stream(stream));
}

} // namespace c12

using namespace c12;

extern "C" {
int c12381_sha3_512_batch(const uint8_t* msgs, size_t msg_len, size_t B, uint8_t* out64) { return hash_host(msgs, msg_len, B, 0, out64); }
int c12381_hash_to_zp_batch(const uint8_t* msgs, size_t msg_len, size_t B, uint8_t* out32) { return hash_host(msgs, msg_len, B, 1, out32); }
int c12381_hash_to_g1_batch(const uint8_t* msgs, size_t msg_len, size_t B, uint8_t* out49) { return hash_host(msgs, msg_len, B, 2, out49); }
int c12381_map_to_g1_batch(const uint8_t* u48, size_t B, uint8_t* out49) { return hash_host(u48, 48, B, 3, out49); }
int c12381_hash_to_g1_batch_dev(const uint8_t* msgs, size_t msg_len, size_t B, uint8_t* out49, void* st) { return hash_dev(msgs, msg_len, B, 2, out49, st); }
int c12381_map_to_g1_batch_dev(const uint8_t* u48, size_t B, uint8_t* out49, void* st) { return hash_dev(u48, 48, B, 3, out49, st); }
int c12381_sha3_512_batch_dev(const uint8_t* msgs, size_t msg_len, size_t B, uint8_t* out64, void* st) { return hash_dev(msgs, msg_len, B, 0, out64, st); }
int c12381_hash_to_zp_batch_dev(const uint8_t* msgs, size_t msg_len, size_t B, uint8_t* out32, void* st) { return hash_dev(msgs, msg_len, B, 1, out32, st); }
}
