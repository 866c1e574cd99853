"""CPU oracle for the 3D U-Net train-step path of torch-em (TEST INFRASTRUCTURE ONLY).

This package restates, in plain fp32 PyTorch / numpy, the arithmetic of the reference functions on the
hot path.  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing under ``torch-em_b200/``
(the product) imports it, and the product has no CPU fallback.

Parity pinning:
  * ``oracle.unet`` / ``oracle.dice`` are pinned against the reference modules themselves
    (``/root/reference/torch_em/model/unet.py``, ``loss/dice.py``, ``loss/wrapper.py``), imported by file
    path in the build container by ``tests/golden/make_golden.py``; the resulting vectors are committed
    under ``tests/golden/`` and re-checked by ``tests/test_oracle.py``.
  * ``oracle.labels.affinity_targets`` is pinned against the brute-force functions the reference's own test
    holds (``test/transform/test_label_transforms.py:5-55``), restated in ``oracle.labels``.
  * ``oracle.labels.boundary_targets``: *parity unpinned* by any reference test (none exists) -- pinned
    against a scipy grey-dilation != grey-erosion restatement of ``skimage.segmentation.find_boundaries``
    (``mode="thick"``), scikit-image itself being absent from this image.
"""
