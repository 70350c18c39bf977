// FieldOps for the field declared by the reference's src/bn256.rs (= BLS12-381 scalar field).
#include "field_ops.cuh"
namespace hodor {
const FieldOps kOpsBlsFr = Ops<BlsFr>::table();
}
