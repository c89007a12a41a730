// tests/plugin/mod_op_device.cu -- device side of the user-defined Ops of mod_op.h: instantiates the
// library's kernel templates over ModOp<T>::apply_device / MyOp<T>::apply_device and registers the
// launchers under their names at load time.  Compiled by nvcc for sm_100a; libsmb200.so is untouched.
#include <smb200_plugin.cuh>
#include "mod_op.h"

SMB_REGISTER_DEVICE_OP("ModOp", int32_t, ModOp<int32_t>);
SMB_REGISTER_DEVICE_OP("ModOp", float, ModOp<float>);
SMB_REGISTER_DEVICE_OP("MyOp", float, MyOp<float>);
SMB_REGISTER_DEVICE_OP("MyOp", double, MyOp<double>);
SMB_REGISTER_DEVICE_OP("MyOp", int32_t, MyOp<int32_t>);
