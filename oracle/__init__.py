"""CPU oracle for the TinyChatEngine quantized-linear / KV-attention hot path.

TEST INFRASTRUCTURE ONLY: importable from tests/, ``__graft_entry__.smoke()`` and bench.py's
``cpu_baseline`` / ``--impl reference`` legs.  The product package (``tinychatengine_b200``) never imports it.

* :mod:`oracle.capi`   -- ctypes bindings to ``libtce_oracle.so`` (plain-C restatement, tce_oracle.c) and to the
  reference's own kernels compiled in place (``oracle/_ref/*.so``, ref_shim.cc).
* :mod:`oracle.quant`  -- numpy port of the reference's offline quantizer formats
  (llm/tools/quantize_methods.py) used to make byte-identical packed inputs.
* :mod:`oracle.sampling` -- numpy restatement of the sampling chain (llm/src/Generate.cc), pinned to the compiled reference.
* :mod:`oracle.llama_ref` -- the Llama step COMPOSED from the pieces above (``llama_forward``), pinned against the reference's whole CPU model
  (Int4LlamaForCausalLM compiled in place); tests/helpers.py::oracle_decode_step runs the same composition for the GPU parity tests.

Parity pinning status: the reference's golden tensors (llm/assets, a download) are absent, so the oracle is
pinned against the reference sources compiled here (tests/test_oracle_golden.py, runs whenever oracle/_ref
exists) and against committed fixtures generated from that build (tests/golden/).
"""
