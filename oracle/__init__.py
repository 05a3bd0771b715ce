"""CPU oracle for the HifiDiff reverse-sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the timed
CPU baseline.  The product (``hifidiff_b200``) never imports this package and
has no CPU fallback.

Contents
--------
``denoiser_ref``   functional PyTorch restatement of the reference's
                   ``models/denoiser`` arithmetic (state_dict in, tensors out).
``cond_ref``       functional restatement of FPG / IDC (the condition-only nets).
``schedulers_ref`` restatement of the diffusers==0.32.2 DDIM / DDPM arithmetic
                   the reference calls (``requirements.txt:6``); the dependency
                   is not vendored in /root/reference and not installed here.
``philox``         numpy Philox4x32-10 + Box-Muller, the counter-based noise the
                   CUDA sampler kernel also implements.
``ref_shim``       (container only) imports the *unmodified* reference modules
                   from /root/reference with a stub ``diffusers.ConfigMixin``.

Parity pinning
--------------
* Model arithmetic: PINNED.  ``tests/golden/make_golden.py`` runs the imported
  reference modules and this restatement on identical weights/inputs, asserts
  agreement, and commits the reference's outputs as fixtures under
  ``tests/golden/``.
* Scheduler arithmetic (diffusers): the reference holds no test or golden vector
  at that boundary and diffusers is absent, so the restatement of the published
  0.32.2 algorithm is PINNED against the dependency's own known-answer tests
  instead: ``tests/test_schedulers.py`` replays the full loops and variance
  checks of diffusers' ``tests/schedulers/test_scheduler_ddim.py`` /
  ``test_scheduler_ddpm.py`` and reproduces their asserted numbers (fixtures and
  constants restated from the published suite; see ``schedulers_ref.py``).
"""
