// FieldOps for the BN254 scalar field (the curve usually called bn256).
#include "field_ops.cuh"
namespace hodor {
const FieldOps kOpsBn254Fr = Ops<Bn254Fr>::table();
}
