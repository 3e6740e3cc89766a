# Convenience targets; the driver's contract is __graft_entry__.build() / smoke() and bench.py.
PY ?= python

build:
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu: build
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu: build
	$(PY) -m pytest tests -x -q -m gpu

smoke: build
	$(PY) -c "import __graft_entry__ as g; g.smoke()"

bench: build
	$(PY) bench.py

bench-reference: build
	$(PY) bench.py --impl reference

bench-configs: build
	$(PY) tools/bench_configs.py

# stream (col, val) + gather only: the bound of a CSR SpMV on the power-law matrix
gather-probe: build
	$(PY) tools/gather_probe.py 10000000

clean:
	$(MAKE) -C repo-8852-ginkgo_b200/csrc clean
	$(MAKE) -C oracle clean
	$(MAKE) -C shim clean
	$(MAKE) -C tools/probe clean

.PHONY: build test-cpu test-gpu smoke bench bench-reference bench-configs gather-probe clean
