# Builds everything in-tree (built artefacts are git-ignored but travel to the GPU box):
#   skid_b200/libskidgpu.so   the CUDA hot path behind the C-ABI of include/skidgpu.h (sm_100a only)
#   host/skid                 flag-compatible C driver (SKID's main.c surface) linked against it
#   oracle/liboracle.so       CPU restatement of the reference algorithm (TEST INFRASTRUCTURE)
#   oracle/_ref/*             the unmodified reference, when /root/reference is present
NVCC ?= /usr/local/cuda/bin/nvcc
CC ?= gcc
NVFLAGS = -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function
CUSRC = $(wildcard skid_b200/csrc/*.cu)
CUOBJ = $(CUSRC:.cu=.o)
CUHDR = $(wildcard skid_b200/csrc/*.cuh) include/skidgpu.h

all: lib host oracle

lib: skid_b200/libskidgpu.so

skid_b200/csrc/%.o: skid_b200/csrc/%.cu $(CUHDR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

# unbinding: the float64 expressions of the energy scan, the centre-of-mass update and SPLINE_POT keep the
# reference's operation order (gcc x86-64, no FMA) - no contraction into DFMA; the explicit fmaf of far_dir stay
skid_b200/csrc/groups.o: skid_b200/csrc/groups.cu $(CUHDR)
	$(NVCC) $(NVFLAGS) -fmad=false -c $< -o $@

skid_b200/libskidgpu.so: $(CUOBJ)
	$(NVCC) -shared -o $@ $(CUOBJ) -gencode arch=compute_100a,code=sm_100a

host: host/skid host/io_harness host/totipnat

HOSTSRC = host/tipsy_io.c host/outputs.c host/cosmo.c host/fastio.c

host/skid: host/skid_main.c $(HOSTSRC) host/skid_host.h include/skidgpu.h skid_b200/libskidgpu.so
	$(CC) -O2 -Wall -Iinclude -pthread -o $@ host/skid_main.c $(HOSTSRC) \
		-Lskid_b200 -lskidgpu -Wl,-rpath,'$$ORIGIN/../skid_b200' -lm

host/totipnat: host/totipnat.c host/fastio.c host/skid_host.h include/skidgpu.h
	$(CC) -O2 -Wall -Iinclude -pthread -o $@ host/totipnat.c host/fastio.c -lm

# host-side I/O without the GPU library: test/benchmark harness for the readers and writers
host/io_harness: tests/io_harness.c $(HOSTSRC) host/skid_host.h include/skidgpu.h
	$(CC) -O2 -Wall -Iinclude -Ihost -pthread -o $@ tests/io_harness.c $(HOSTSRC) -lm

oracle: oracle/liboracle.so ref

oracle/liboracle.so: oracle/skid_oracle.c oracle/skid_oracle.h
	$(CC) -O2 -fPIC -shared -Wall -ffp-contract=off -o $@ oracle/skid_oracle.c -lm

ref:
	./oracle/build_ref.sh

clean:
	rm -f skid_b200/csrc/*.o skid_b200/libskidgpu.so host/skid host/io_harness host/totipnat oracle/liboracle.so

.PHONY: all lib host oracle ref clean
