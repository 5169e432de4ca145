# Convenience targets (the driver uses __graft_entry__.build(), bench.py and pytest directly).
PY ?= python
.PHONY: build test-cpu test-gpu test-emulated bench clean
build:
	$(PY) -c "import __graft_entry__ as g; g.build()"
test-cpu: build
	$(PY) -m pytest tests -q -m "not gpu"
test-gpu: build
	$(PY) -m pytest tests -q -m gpu
# the -m gpu suite on the kernels' own source executed by the SIMT emulator (no GPU needed, ~7 min)
test-emulated: build
	$(PY) -c "import sys; sys.path.insert(0, 'tests/simt_emu'); import build_emu; build_emu.build()"
	B200NB_LIB=$(CURDIR)/tests/simt_emu/_build/libb200nb_emu.so $(PY) -m pytest tests -q -m gpu
bench: build
	$(PY) bench.py
clean:
	$(MAKE) -C deseq2_b200/csrc clean
	$(MAKE) -C oracle clean || true
	rm -rf tests/simt_emu/_build
