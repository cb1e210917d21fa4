# Convenience targets (everything real lives in focusflow_official_b200/csrc/Makefile and oracle/Makefile).
PY ?= python

build:            ## libffcorr.so (sm_100a), the oracle's C restatement and, where /root/reference exists, the reference's PWC kernels
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu: build   ## oracle vs golden vectors, host logic, C-ABI exports; no GPU needed
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu: build   ## parity through the C ABI; run on a B200
	$(PY) -m pytest tests -x -q -m gpu

smoke: build
	$(PY) __graft_entry__.py smoke

bench: build      ## one JSON line; add GPUS=N for torchrun
	$(PY) bench.py --steps 5 --warmup 3

clean:
	$(MAKE) -C focusflow_official_b200/csrc clean
	$(MAKE) -C oracle clean

.PHONY: build test-cpu test-gpu smoke bench clean
