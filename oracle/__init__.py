"""TEST INFRASTRUCTURE ONLY -- CPU/PyTorch fp32 restatement of the SYNT_ISIC hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and there only as the checker (or as the timed CPU baseline),
never as the thing shipped.  The product package ``synt_isic_b200`` never imports
this package and fails loudly when its CUDA library is missing.

What is restated (reference paths relative to /root/reference):

* ``unet2d.py``   -- diffusers ``UNet2DModel`` for the constructor arguments at
  ``core/generator/model_manager.py:175-194``.
* ``ddpm.py``     -- diffusers ``DDPMScheduler`` as configured at
  ``core/generator/model_manager.py:199-209`` (Appendix A.3).
* ``classifier.py`` -- ``MelanomaClassifierAdaptive`` (``xai/XAI.py:357-471``) on the
  REAL torchvision ``resnet18``.
* ``xai.py``      -- ``compute_time_shap`` (``xai/XAI.py:1179-1234``), patch-SHAP
  (``:1111-1177``), interventions (``:1454-1597``) and CFI (``:1600-1700``).

PARITY PIN STATUS
-----------------
``diffusers`` is an un-vendored, unpinned (``>=0.21.0``, requirements.txt:6)
third-party dependency that is NOT installed in this image, and the reference has no
tests or golden vectors for this path.  The UNet/scheduler oracle is therefore
**parity unpinned** against real diffusers; it is pinned only by the indirect
known-answer checks of SURVEY.md section 8(c) (parameter count 25,304,963 from the
shipped checkpoint sizes, strict state_dict key scheme, timestep tables, schedule
constants, seed algebra).  The classifier oracle runs the real torchvision code and the
real ``F.interpolate`` call, so it is pinned by construction.
"""
