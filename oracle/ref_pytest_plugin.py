"""pytest plugin that lets the reference's OWN test-suite run in this image:

    cd /tmp && python -m pytest -p oracle.ref_pytest_plugin -p no:cacheprovider --rootdir /tmp -c /dev/null \\
        -m "not slow" /root/reference/tests            # 23 passed, 6 deselected

TEST INFRASTRUCTURE ONLY.  It registers the TF / TFP stand-ins of ``oracle/tf_shim.py`` and the
``estimators`` / ``evaluation`` package objects BEFORE collection, so the unmodified test files import the
unmodified estimator, layer, flow and scorer modules from /root/reference (training runs on the stand-ins'
minimal Keras loop: mini-batch Adam through torch autograd).  The
one adaptation: ``KMeans(n_jobs=-2)`` (DistributionLayers.py:162) no longer exists in scikit-learn, so
the ``KMeans`` name inside the reference module is wrapped to drop ``n_jobs``.
``tests/test_reference_run.py::test_reference_own_unit_tests_pass_on_the_stand_ins`` runs this.
"""
import sys

from oracle import tf_shim

sys.dont_write_bytecode = True  # nothing is written into the read-only reference tree
_PKG = tf_shim.load_reference_package()
_DL = sys.modules["estimators.DistributionLayers"]

from sklearn.cluster import KMeans as _KMeans  # noqa: E402

_DL.KMeans = lambda n_clusters, n_jobs=None: _KMeans(n_clusters=n_clusters, n_init=10, random_state=22)
