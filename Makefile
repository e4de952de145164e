# Build the sm_100a shared library in-tree (it travels to the GPU box with the snapshot).
NVCC ?= nvcc
NVFLAGS = -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
SRC = qm_door_b200/csrc/qmb200.cu $(wildcard qm_door_b200/csrc/*.cpp)
HDR = $(wildcard qm_door_b200/csrc/*.h) include/qmb200.h

all: qm_door_b200/libqmb200.so oracle/cport/libcport.so

qm_door_b200/libqmb200.so: $(SRC) $(HDR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRC)

oracle/cport/libcport.so: oracle/cport/cport.cpp $(HDR)
	g++ -O3 -march=x86-64-v3 -std=c++17 -shared -fPIC -pthread -o $@ oracle/cport/cport.cpp

clean:
	rm -f qm_door_b200/libqmb200.so oracle/cport/libcport.so

# development build with per-phase cycle counters (tools/phase_timing.py)
dbg: $(SRC) $(HDR)
	$(NVCC) $(NVFLAGS) -DQM_PHASE_TIMING -shared -o qm_door_b200/libqmb200_dbg.so $(SRC)
