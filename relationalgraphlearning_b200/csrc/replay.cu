// GPU-resident replay memory for the MPRL trainer (crowd_nav/utils/memory.py:4-28 stores python tuples of tiny tensors and
// torch's DataLoader collates 6 x batch_size of them per minibatch, crowd_nav/utils/trainer.py:66-67,113-114).
// Here a transition is ONE contiguous record in HBM
//   [ robot 9 | humans 5*Nh | value 1 | reward 1 | next_robot 9 | next_humans 5*Nh ]      rec = 20 + 10*Nh floats
// and a minibatch is one gather launch that writes the six batch tensors the trainer consumes.
#include "kernels.h"

namespace rgl {

// one thread per (sample b, float f of the record): reads store[idx[b]][f] (coalesced inside a record), writes the field's
// batch tensor (coalesced inside a sample)
__global__ void replay_gather_kernel(const float* __restrict__ store, const long long* __restrict__ idx, int B, int Nh, int rec,
                                     float* __restrict__ robot, float* __restrict__ humans, float* __restrict__ value,
                                     float* __restrict__ reward, float* __restrict__ next_robot, float* __restrict__ next_humans) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * rec) return;
    const int b = (int)(t / rec), f = (int)(t - (long long)b * rec);
    const float v = store[idx[b] * rec + f];
    const int hw = HD * Nh;
    if (f < RD) robot[(size_t)b * RD + f] = v;
    else if (f < RD + hw) humans[(size_t)b * hw + f - RD] = v;
    else if (f == RD + hw) value[b] = v;
    else if (f == RD + hw + 1) reward[b] = v;
    else if (f < 2 * RD + hw + 2) next_robot[(size_t)b * RD + f - (RD + hw + 2)] = v;
    else next_humans[(size_t)b * hw + f - (2 * RD + hw + 2)] = v;
}

// one thread per float of the record: writes one transition (six device tensors) into its slot
__global__ void replay_push_kernel(float* __restrict__ store, long long slot, int Nh, int rec, const float* __restrict__ robot,
                                   const float* __restrict__ humans, const float* __restrict__ value, const float* __restrict__ reward,
                                   const float* __restrict__ next_robot, const float* __restrict__ next_humans) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= rec) return;
    const int hw = HD * Nh;
    float v;
    if (f < RD) v = robot[f];
    else if (f < RD + hw) v = humans[f - RD];
    else if (f == RD + hw) v = value[0];
    else if (f == RD + hw + 1) v = reward[0];
    else if (f < 2 * RD + hw + 2) v = next_robot[f - (RD + hw + 2)];
    else v = next_humans[f - (2 * RD + hw + 2)];
    store[slot * rec + f] = v;
}

cudaError_t run_replay_gather(const float* store, const long long* idx, int B, int Nh, float* robot, float* humans, float* value,
                              float* reward, float* next_robot, float* next_humans, cudaStream_t st) {
    const int rec = 2 * RD + 2 * HD * Nh + 2;
    const long long total = (long long)B * rec;
    replay_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(store, idx, B, Nh, rec, robot, humans, value, reward, next_robot,
                                                                          next_humans);
    return cudaGetLastError();
}

cudaError_t run_replay_push(float* store, long long slot, int Nh, const float* robot, const float* humans, const float* value,
                            const float* reward, const float* next_robot, const float* next_humans, cudaStream_t st) {
    const int rec = 2 * RD + 2 * HD * Nh + 2;
    replay_push_kernel<<<(rec + 127) / 128, 128, 0, st>>>(store, slot, Nh, rec, robot, humans, value, reward, next_robot, next_humans);
    return cudaGetLastError();
}

}  // namespace rgl
