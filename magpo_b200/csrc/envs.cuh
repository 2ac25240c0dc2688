// Internal launchers shared between the env kernels and the rollout sequencer.
#pragma once
#include "common.cuh"

namespace magpo {

// vmap(env.step) for CoordSum; done_out [B] (optional) receives timestep.last().
int coordsum_step_launch(cudaStream_t s, const MagpoCoordSumCfg* cfg, int B, const int32_t* action,
                         MagpoCoordSumState st, MagpoTimeStep ts, uint8_t* done_out);
// vmap(env.step) for LevelBasedForaging.
int lbf_step_launch(cudaStream_t s, const MagpoLbfCfg* cfg, int B, const int32_t* action, MagpoLbfState st,
                    MagpoTimeStep ts, uint8_t* done_out);

// vmap(env.step) for RobotWarehouse.
int rware_step_launch(cudaStream_t s, const MagpoRwareCfg* cfg, int B, const int32_t* action, MagpoRwareState st,
                      MagpoTimeStep ts, uint8_t* done_out);

}  // namespace magpo
