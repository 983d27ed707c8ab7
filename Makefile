# neutral with the b200 kernel set: `make KERNELS=b200` builds neutral.b200, the reference's own
# C driver (main.c + neutral_data.c, compiled UNMODIFIED from where they lie under $(REF))
# linked against libneutral_b200.so - the drop-in next to omp3/oacc/cuda that the reference's
# Makefile:2,83-90 selects with the same variable. `make KERNELS=omp3` builds the reference's
# own omp3 set the same way (that is the oracle/baseline build, see oracle/Makefile).
#
# Knobs mirror the reference Makefile:2-6,45-51: KERNELS, COMPILER (GCC only here), DEBUG,
# OPTIONS; NGPUS and DECK belong to the `run` target (`make run DECK=split NGPUS=8` shards the
# bank over 8 GPUs of the box: the library reads NB200_NGPUS, the driver is unchanged). The absent `arch` parent project is served by archlite/ (SURVEY.md appendix A).
# Outputs land in build/ (git-ignored; it travels to the GPU box through gpurun):
#   build/run/neutral/neutral.$(KERNELS)   run it from build/run/neutral, like the reference:
#       cd build/run/neutral && ./neutral.b200 problems/csp.params
KERNELS  ?= b200
COMPILER ?= GCC
DEBUG    ?= no
REF      ?= /root/reference
OPTIONS  += -g -DENABLE_PROFILING -D__STDC_CONSTANT_MACROS

ROOT     := $(dir $(abspath $(lastword $(MAKEFILE_LIST))))
SHIM     := $(ROOT)archlite
RUN      := $(ROOT)build/run/neutral
CC       := gcc
NVCC     ?= nvcc
# reference CFLAGS_GCC (Makefile:13) with the two parity deviations documented in
# oracle/Makefile: -ffp-contract=off and a portable -march.
CFLAGS_GCC := -O3 -std=gnu99 -fopenmp -march=x86-64-v3 -ffp-contract=off -Wall \
              -Wno-unused-variable -Wno-unused-but-set-variable -Wno-format -Wno-unused-result
ifeq ($(DEBUG), yes)
  OPTIONS += -O0 -DDEBUG
endif
INC      := -I$(SHIM)/a -I$(SHIM)/a/b -I$(SHIM)
B200LIB  := $(ROOT)neutral_b200/libneutral_b200.so

NGPUS    ?= 1
DECK     ?= csp

.PHONY: neutral lib clean rundir run
neutral: $(RUN)/neutral.$(KERNELS)

run: neutral
	cd $(RUN) && NB200_NGPUS=$(NGPUS) ./neutral.$(KERNELS) problems/$(DECK).params

lib $(B200LIB):
	python -m neutral_b200.build

rundir:
	@mkdir -p $(RUN)
	@cp $(SHIM)/arch.params $(ROOT)build/run/arch.params
	@test -e $(RUN)/elastic_scatter.cs || install -m 0644 $(REF)/elastic_scatter.cs $(REF)/capture.cs $(RUN)/
	@ln -sfn ../../../problems $(RUN)/problems

ifeq ($(KERNELS), b200)
# GPU kernel sets of the reference use the SoA Particle (reference Makefile:60-71).
OPTIONS += -DSoA
$(RUN)/neutral.b200: $(REF)/main.c $(REF)/neutral_data.c $(SHIM)/archlite.c $(B200LIB) | rundir
	$(CC) $(CFLAGS_GCC) $(OPTIONS) $(INC) $(REF)/main.c $(REF)/neutral_data.c $(SHIM)/archlite.c \
	  -L$(ROOT)neutral_b200 -lneutral_b200 -Wl,-rpath,'$$ORIGIN/../../../neutral_b200' -lm -o $@
else ifeq ($(KERNELS), omp3)
$(RUN)/neutral.omp3: $(REF)/main.c $(REF)/neutral_data.c $(REF)/omp3/neutral.c $(SHIM)/archlite.c $(SHIM)/alloc_host.c | rundir
	$(CC) $(CFLAGS_GCC) $(OPTIONS) $(INC) $^ -lm -o $@
else
$(error KERNELS must be b200 or omp3 (the other reference sets need compilers/libraries this image lacks))
endif

clean:
	rm -rf $(ROOT)build
