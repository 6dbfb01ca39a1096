"""pytest plugin that lets the reference's OWN unit tests for the path run in this image:

    python -m pytest -p oracle.ref_pytest_plugin -p no:cacheprovider \\
        /root/reference/tests/test_flows.py /root/reference/tests/test_distribution_layers.py

TEST INFRASTRUCTURE ONLY.  It registers the TF / TFP stand-ins of ``oracle/tf_shim.py`` and the
``estimators`` package object BEFORE collection, so the unmodified test files import the unmodified
``estimators/normalizing_flows`` and ``estimators/DistributionLayers.py`` from /root/reference.  The
one adaptation: ``KMeans(n_jobs=-2)`` (DistributionLayers.py:162) no longer exists in scikit-learn, so
the ``KMeans`` name inside the reference module is wrapped to drop ``n_jobs``.
``tests/test_reference_run.py::test_reference_own_unit_tests_pass_on_the_stand_ins`` runs this.
"""
import sys

from oracle import tf_shim

sys.dont_write_bytecode = True  # nothing is written into the read-only reference tree
_FLOWS, _DL = tf_shim.load_reference()

from sklearn.cluster import KMeans as _KMeans  # noqa: E402

_DL.KMeans = lambda n_clusters, n_jobs=None: _KMeans(n_clusters=n_clusters, n_init=10, random_state=22)
