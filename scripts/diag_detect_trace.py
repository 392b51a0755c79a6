"""Host-side timeline of detectTriangles around its synchronisations at S3 ticks 70..73 (PIES_B200_DETECT_TRACE=1)."""
import os, sys
sys.path.insert(0, ".")
os.environ.setdefault("PIES_B200_DETECT_TRACE", "1")
import pies_b200 as pb
from pies_b200 import scenes
s = pb.Solver(**scenes.S3_OPTIONS)
scenes.build_s3(s, int(os.environ.get("BODIES", "20834")))
s.setTuning(profilePhases=True)
for t in range(1, 74):
    if t == 70:
        sys.stderr.write("---- tick 70\n")
    s.tick()
    if t >= 70:
        st = s.stats()
        sys.stderr.write("tick %d ms: tick %.3f detect %.3f contact %.3f global %.3f local %.3f\n" % (t, st.msTick, st.msDetect, st.msContact, st.msGlobal, st.msLocal))
