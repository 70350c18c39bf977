// FieldOps for the 252-bit Stark prime of the reference's experiments (src/experiments/mod.rs:18-21).
#include "field_ops.cuh"
namespace hodor {
const FieldOps kOpsStark252 = Ops<Stark252>::table();
}
