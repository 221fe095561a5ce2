"""TEST INFRASTRUCTURE ONLY — the synthetic weight / input generators live in tools/synth.py (bench.py's product arm uses
them too, and must not import anything from oracle/); re-exported here for the oracle-side scripts and the tests."""
from tools.synth import *  # noqa: F401,F403
from tools.synth import Gen  # noqa: F401
