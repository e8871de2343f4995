# Convenience targets; the Python entry points do the same work (__graft_entry__.build()).
NVCC ?= nvcc
LIB  := cudavideostream_b200/libcvs_b200.so
CSRC := cudavideostream_b200/csrc

all: lib oracle

lib: $(LIB)
$(LIB): $(wildcard $(CSRC)/*.cu $(CSRC)/*.cuh include/*.h include/*.hpp)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden \
	    -shared -cudart shared -I include -o $@ $(CSRC)/cvs_api.cu $(CSRC)/cvs_shim.cu -ldl

oracle:
	$(MAKE) -C oracle -s all

# a reference-style C++ host (plain g++, no CUDA headers) linked against the library
shim_host: lib tests/host/shim_host.cpp
	g++ -std=c++11 -O1 -I include tests/host/shim_host.cpp -o $@ -Lcudavideostream_b200 -l:libcvs_b200.so \
	    -Wl,-rpath,$(abspath cudavideostream_b200)

test:
	python -m pytest tests -q -m "not gpu"

clean:
	rm -f $(LIB) shim_host
	$(MAKE) -C oracle -s clean

.PHONY: all lib oracle test clean
