"""CPU oracle for the waveform-generation hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker or the timed CPU
baseline — never as the thing shipped.  The product package
(``megatts2_hierspeechpp_b200``) does not import this package and has no CPU
fallback.

Contents
--------
``functional``   op-for-op torch-CPU fp32 restatement of the reference modules
                 (state_dict in, tensors out; no nn.Module).
``closed_form``  independent numpy fp64 closed forms (Activation1d polyphase
                 form, interpolation index tables, padding/phase tables).
``synth``        seeded synthetic checkpoints / inputs.
``refload``      import shim for the real reference under /root/reference
                 (import stubs for timm / monotonic_align).  Only usable in
                 the authoring container; used to PIN the oracle and to
                 generate ``tests/golden``.

Parity pin status: the reference ships no tests or golden vectors for this
path.  The oracle is pinned by running the reference's own
modules beside it in the authoring container (tests/test_oracle_vs_reference.py,
skipped where /root/reference is absent) and by the committed fixtures under
tests/golden/ that tests/gen_golden.py produced from the reference.
"""
