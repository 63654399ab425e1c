#!/bin/bash
# Kernel-variant experiments: benchmarks/build_variant.sh UNIT NAME "-DFLAG=.. .."
#   recompiles only csrc/UNIT.cu with the flags and links build/variants/libUNIT_NAME.so against the in-tree objects
#   (build() first); run a benchmark or the tests against it with CRB_LIB=$PWD/build/variants/libUNIT_NAME.so.
# Flags used in round 2 (all measured, see DESIGN.md 4 "RK45" / benchmarks/experiments/run_r2w.sh .. run_r2ah.sh):
#   crb_rk45: -DCRB_RK45_ROLLED=0|1  -DCRB_RK45_WPB=1|2|3|4|8  -DCRB_RK45_LOCKSTEP=1  -DCRB_RK45_QVS=1
#             -DCRB_RK45_MAXNREG=168|184|200  -DCRB_RK45_UMS_ALL=0|1      crb_rk4: -DCRB_RK4_UMS_ALL=1
set -e
cd "$(dirname "$0")/.."
UNIT=$1; NAME=$2; shift; shift
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Iinclude -Xcompiler -fPIC $@ -c continuum_robot_b200/csrc/$UNIT.cu -o build/variants/${UNIT}_$NAME.o
OBJS=$(ls continuum_robot_b200/csrc/*.o | grep -v "/$UNIT.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib${UNIT}_$NAME.so $OBJS build/variants/${UNIT}_$NAME.o
echo built $UNIT $NAME
