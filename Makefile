# Top-level build.
#   make            CUDA C-ABI library (nvcc, sm_100a) + host C libraries + oracle checkers
#   make emu        CPU-emulator builds used by the CPU-only unit tests (test infrastructure)
#   make stream     host/bchstream: streaming decoder CLI on the drop-in library
#   make demos      build the reference's own fftdemo/bchdemo sources UNCHANGED against our
#                   libraries (needs /root/reference; drop-in check)
ROOT := $(dir $(abspath $(lastword $(MAKEFILE_LIST))))
CC   ?= gcc
CSRC := $(ROOT)bch_fft_b200/csrc
HOSTC := $(ROOT)host/base.c $(ROOT)host/context.c $(ROOT)host/fft.c $(ROOT)host/bch.c
HFLAGS := -std=gnu99 -O2 -fPIC -Wall -Wextra -I$(ROOT)include -I$(ROOT)host
REF  ?= /root/reference

all: cuda host oracle stream
# where the reference tree exists (the build container) its demo sources are rebuilt against the
# fresh libraries on every build, so the binaries that travel to the GPU box are never stale
ifneq ($(wildcard $(REF)/bch/main.c),)
all: demos
endif
cuda:
	$(MAKE) -C $(CSRC) all
emu: oracle
	$(MAKE) -C $(CSRC) emu
	$(CC) $(HFLAGS) -shared -o $(ROOT)host/libbchfft_host_emu.so $(HOSTC) \
	    -L$(CSRC) -lbchfft_emu -Wl,-rpath,'$$ORIGIN/../bch_fft_b200/csrc' -lpthread
host: cuda
	$(CC) $(HFLAGS) -shared -o $(ROOT)host/libbchfft_host.so $(HOSTC) \
	    -L$(CSRC) -lbchfft_cuda -Wl,-rpath,'$$ORIGIN/../bch_fft_b200/csrc' -lpthread
	@mkdir -p $(ROOT)host/build
	$(CC) $(HFLAGS) -c $(ROOT)host/base.c -o $(ROOT)host/build/base.o
	$(CC) $(HFLAGS) -c $(ROOT)host/context.c -o $(ROOT)host/build/context.o
	$(CC) $(HFLAGS) -c $(ROOT)host/fft.c -o $(ROOT)host/build/fft.o
	$(CC) $(HFLAGS) -c $(ROOT)host/bch.c -o $(ROOT)host/build/bch.o
	ar rcs $(ROOT)host/libbase.a $(ROOT)host/build/base.o $(ROOT)host/build/context.o
	ar rcs $(ROOT)host/libfft.a $(ROOT)host/build/fft.o
	ar rcs $(ROOT)host/libbch.a $(ROOT)host/build/bch.o
oracle:
	$(MAKE) -C $(ROOT)oracle
# the reference's demo sources, unmodified, against our drop-in libraries
demos: host
	$(CC) -std=gnu99 -O2 -I$(ROOT)include -o $(ROOT)host/fftdemo $(REF)/fft/test.c \
	    -L$(ROOT)host -lfft -lbase -L$(CSRC) -lbchfft_cuda -Wl,-rpath,$(CSRC) -lpthread -lm
	$(CC) -std=gnu99 -O2 -I$(ROOT)include -o $(ROOT)host/bchdemo $(REF)/bch/main.c \
	    -L$(ROOT)host -lbch -lfft -lbase -L$(CSRC) -lbchfft_cuda -Wl,-rpath,$(CSRC) -lpthread -lm
demos-emu: emu
	$(CC) -std=gnu99 -O2 -I$(ROOT)include -o $(ROOT)host/fftdemo_emu $(REF)/fft/test.c \
	    -L$(ROOT)host -lbchfft_host_emu -Wl,-rpath,$(ROOT)host -lm
	$(CC) -std=gnu99 -O2 -I$(ROOT)include -o $(ROOT)host/bchdemo_emu $(REF)/bch/main.c \
	    -L$(ROOT)host -lbchfft_host_emu -Wl,-rpath,$(ROOT)host -lm
# streaming decoder CLI: packed codewords stdin -> stdout through bch_decode_batch
stream: host
	$(CC) -std=gnu99 -O2 -Wall -Wextra -I$(ROOT)include -o $(ROOT)host/bchstream $(ROOT)host/bchstream.c \
	    -L$(ROOT)host -lbchfft_host -L$(CSRC) -lbchfft_cuda -Wl,-rpath,'$$ORIGIN' \
	    -Wl,-rpath,'$$ORIGIN/../bch_fft_b200/csrc' -lpthread -lm
stream-emu: emu
	$(CC) -std=gnu99 -O2 -Wall -Wextra -I$(ROOT)include -o $(ROOT)host/bchstream_emu $(ROOT)host/bchstream.c \
	    -L$(ROOT)host -lbchfft_host_emu -L$(CSRC) -lbchfft_emu -Wl,-rpath,'$$ORIGIN' \
	    -Wl,-rpath,'$$ORIGIN/../bch_fft_b200/csrc' -lpthread -lm
clean:
	$(MAKE) -C $(CSRC) clean
	$(MAKE) -C $(ROOT)oracle clean
	rm -rf $(ROOT)host/build $(ROOT)host/*.so $(ROOT)host/*.a $(ROOT)host/*demo*
.PHONY: all cuda emu host oracle demos demos-emu stream stream-emu clean
