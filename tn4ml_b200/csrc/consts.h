// Compile-time geometry shared by the host-side planner (api.cu) and the kernel translation units.
#pragma once
#include <stdint.h>

namespace tnb {

constexpr int kRN = 4;  // SIMT family: register tile of RB rows (template parameter, 4 or 8) x kRN columns per thread

// tensor-core chain kernel (mps_tc_chain.cuh)
constexpr int kTcRows = 128;   // samples per CTA == TMEM lanes
constexpr int kTcKA = 14;      // A operand pre-scale: |v| < 1  ->  |v * 2^14| < 2^14 (fp16 max 65504)
constexpr int kTcEpiWarps = 16;                      // 4 lane quarters x 4 column groups
constexpr int kTcMmaWarp = kTcEpiWarps;              // + 1: weight producer warp, + 2: image writer warp
constexpr int kTcThreads = (kTcEpiWarps + 3) * 32;
constexpr int kTcMaxNH = 16;                         // column chunks per step (chi / 16 at most)
constexpr int kTcSmemMax = 232448;                   // opt-in dynamic shared memory of the tensor kernels

// tensor-core weight-gradient GEMM (mps_tc.cuh)
#ifndef TNB_DA_KC
#define TNB_DA_KC 32
#endif
constexpr int kDaKC = TNB_DA_KC;                  // samples per pipeline stage (a multiple of the MMA's K = 16)
constexpr int kDaProducers = 512;                 // threads (16 warps)
constexpr int kDaThreads = kDaProducers + 64;     // + MMA warp + bulk-copy warp
constexpr int kDaMaxCpg = 8;                      // chi >= 32
// U operand of a stage: a "virtual bond image" of 256 rows, [kDaKC/8 sample blocks][plane hi|lo][32 row octets][8 samples][16 B]
constexpr uint32_t kDaUPlane = 32u * 128u;        // 4 KB
constexpr uint32_t kDaUBlock = 2u * kDaUPlane;    // 8 KB: K-block (8 samples) stride
constexpr uint32_t kDaU = (kDaKC / 8) * kDaUBlock;  // 16 KB
constexpr uint32_t kDaBMax = 256u * 4u * kDaKC;   // V region: a raw image chunk of kDaKC samples
constexpr uint32_t kDaStage = kDaU + kDaBMax;     // 64 KB (32 samples): 3 stages

constexpr int kSimtSmemMax = 200 * 1024;          // opt-in ceiling of the SIMT sweep / class kernels
constexpr int kSmpoSmemMax = 210 * 1024;          // ... of the SMPO kernels

}  // namespace tnb
