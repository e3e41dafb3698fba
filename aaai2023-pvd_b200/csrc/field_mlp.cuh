// field_mlp.cuh -- constants shared by the fused NeRF-MLP forward (field_mlp.cu) and backward (field_mlp_bwd.cu): packed weight
// blob layout, the streamed-piece schedule, and the per-tile layout of what the training forward saves for the backward.
#pragma once
#include "common.cuh"
#include "field_tail.cuh"

namespace pvd {

constexpr uint32_t kMlpChunkBytes = 256 * 64 * 2;   // one K-chunk of a 256-row weight matrix
constexpr uint32_t kMlpChunks = 26;                 // 1 (L0) + 3*4 (L1-3) + 1+4 (L4: in_pts part, hidden part) + 2*4 (L5-6)
constexpr uint32_t kMlpL7Bytes = 4 * (32 * 64 * 2); // 256 -> 28 as four [32 x 64] chunks
constexpr uint32_t kMlpBiasOff = kMlpChunks * kMlpChunkBytes + kMlpL7Bytes;
constexpr uint32_t kMlpBiasBytes = 8 * 256 * 4;
static_assert(kMlpBiasOff + kMlpBiasBytes == PVD_MLP_WBLOB_BYTES, "mlp blob size");

struct MlpArgs {
    const uint8_t* wblob;      // this CTA's replica is selected in the kernel: wblob + (blockIdx.x % replicas) * PVD_MLP_WBLOB_BYTES
    uint32_t replicas;
    const uint8_t* tail_blob;  // PVD_FIELD_WBLOB_BYTES (sigma_net / color_net)
    float clip_min, clip_max, density_scale;
    uint32_t diag;   // timing diagnostics only (PVD_MLP_DIAG): 1 no bias loads, 2 no operand stores, 4 no TMEM loads, 8 no PE
};

// layer schedule: for each of the 26 streamed chunks, which layer it belongs to and where its A operand comes from
struct ChunkDesc {
    uint8_t layer;     // 0..6
    uint8_t from_x0;   // A = X0 tile (in_pts) instead of the activation tile
    uint8_t k_chunk;   // which 64-column slice of the activation tile
    uint8_t last;      // last chunk of its layer
};
static __constant__ ChunkDesc kSchedule[kMlpChunks] = {
    {0, 1, 0, 1},
    {1, 0, 0, 0}, {1, 0, 1, 0}, {1, 0, 2, 0}, {1, 0, 3, 1},
    {2, 0, 0, 0}, {2, 0, 1, 0}, {2, 0, 2, 0}, {2, 0, 3, 1},
    {3, 0, 0, 0}, {3, 0, 1, 0}, {3, 0, 2, 0}, {3, 0, 3, 1},
    {4, 1, 0, 0}, {4, 0, 0, 0}, {4, 0, 1, 0}, {4, 0, 2, 0}, {4, 0, 3, 1},
    {5, 0, 0, 0}, {5, 0, 1, 0}, {5, 0, 2, 0}, {5, 0, 3, 1},
    {6, 0, 0, 0}, {6, 0, 1, 0}, {6, 0, 2, 0}, {6, 0, 3, 1},
};


constexpr uint32_t kPiece = 16384;                    // bytes per streamed weight piece: [256 x 32] fp16, half of a chunk
constexpr uint32_t kPieces = 2 * kMlpChunks + 1;      // 52 half-chunks of layers 0-6 + layer 7
constexpr uint32_t kStages = 4;

// What the TRAINING forward saves per 128-sample tile (fp16 operand tiles in the chunk layout, i.e. exactly the bytes the tensor
// core consumed): the PE tile [128 x 64] and act_1..act_7 = the ReLU outputs of layers 0..6 [128 x 256].
constexpr uint32_t kSavePe = 0;
constexpr uint32_t kSaveAct = 16384;                  // act_l at kSaveAct + (l - 1) * 65536
constexpr uint32_t kSaveTileBytes = PVD_MLP_SAVE_TILE_BYTES;
static_assert(kSaveTileBytes == 16384 + 7 * 65536, "save tile");
// What the trunk backward (data gradients) leaves per tile for the weight-gradient kernel: G7 = d(x28) as a [128 x 32] chunk tile
// and G_0..G_6 = the gradients w.r.t. the pre-activations of layers 0..6 [128 x 256].
constexpr uint32_t kGradG7 = 0;
constexpr uint32_t kGradG = 8192;                     // G_l at kGradG + l * 65536
constexpr uint32_t kGradTileBytes = PVD_MLP_GRAD_TILE_BYTES;
static_assert(kGradTileBytes == 8192 + 7 * 65536, "grad tile");

struct V2Wait {  // bounded waits that stop costing time after the first failure (a wrong barrier must not hang the GPU)
    int32_t* status;
    bool dead = false;
    __device__ __forceinline__ void operator()(uint64_t* bar, uint32_t parity) {
        if (!tc5::mbar_wait(bar, parity, dead ? 1u : (1u << 22))) {
            dead = true;
            atomicExch(status, 2);
        }
    }
};

}  // namespace pvd
