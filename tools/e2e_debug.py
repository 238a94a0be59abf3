"""Debug helper: run the e2e integrator comparison verbosely (prints both result dicts)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
oracle.reference()
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_integrator_e2e as t
import logging
logging.disable(logging.CRITICAL)
want = t.run_once(t.numpy_loglike)
print("REF ", want)
import ultranest_b200
ultranest_b200.install(force=True)
got = t.run_once(t.numpy_loglike)
print("OURS", got)
