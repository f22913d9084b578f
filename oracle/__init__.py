"""TEST INFRASTRUCTURE — not product code.

``oracle/`` holds CPU restatements of the reference's quantized forward path:

* ``ref_harness``  – imports the *unmodified* reference from /root/reference (only exists in the
                     build container) to generate golden vectors; never used on the GPU box.
* ``fakequant``    – torch-CPU restatement of the reference's fp32/fp64 "fake-quant" forward
                     (what the reference actually executes); also the CPU baseline ("port").
* ``int_ref``      – exact integer restatement (numpy int64 / exact fp64 GEMM) = the arbiter the
                     CUDA kernels are compared with bit-for-bit.
* ``int_ref.c``    – plain-C restatement of the same integer arithmetic (built into oracle/_build).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import anything from here.  The product package ``hawq_b200`` never does.
"""
