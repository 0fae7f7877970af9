"""TEST INFRASTRUCTURE ONLY.

oracle/ holds (a) a CPU restatement of the reference rasterizer algorithm (splat_oracle.c +
oracle.py) and (b) the recipe that compiles the *unmodified* reference CUDA sources into
oracle/_ref/ (build_ref.py).  Nothing under saro_gs_b200/ imports this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / reference legs do, and only as the
checker or the timed baseline — never as the product path.
"""
