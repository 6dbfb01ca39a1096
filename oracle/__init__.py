"""CPU oracle for the flow-chain / mixture-head log-likelihood hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline.  The product path (``normalizingflownetwork_b200``) never
imports this package and has no CPU fallback.

PARITY UNPINNED: the reference (siboehm/NormalizingFlowNetwork) is TensorFlow +
TensorFlow-Probability; neither is installed or installable in this image and the
reference's tests hold no golden vectors for this path.  The oracle
is therefore a literal restatement of the reference's in-repo formulas plus the
documented TFP glue semantics, pinned three ways instead:
  * ``oracle/flow_oracle.py``   literal op-for-op torch-CPU restatement (fp64/fp32),
                                gradients from torch autograd;
  * ``oracle/analytic_np.py``   independent NumPy float64 closed-form forward+backward;
  * ``oracle/known_answers_mp.py`` independent 50-digit mpmath scalar evaluation of
                                the reference's own test inputs (``tf.ones`` cases),
frozen into ``tests/golden/*.json`` by ``oracle/make_golden.py`` and cross-checked
against the values derived independently in SURVEY.md App. A.7.

What IS pinned by execution: ``oracle/tf_shim.py`` provides torch-CPU float64 stand-ins for
the ~20 TF ops and 7 TFP glue classes the path touches, so that the reference's own
``estimators/normalizing_flows/*.py`` and ``estimators/DistributionLayers.py`` are imported
UNMODIFIED from /root/reference and run on the golden inputs
(``oracle/make_reference_run.py`` -> ``tests/golden/reference_run.json``).  Their log-probs
and the autograd gradients through them equal the oracle's to 1e-12 / 1e-9 on all 24 chain,
6 MDN and 2 KMN cases and on the reference's ``tf.ones`` test inputs.  That pins every
formula, constant, slice and ordering decision made in the reference's repository; TFP's own
glue arithmetic (restated in the shim from its documented semantics) and float32 rounding
remain unpinned.
"""
