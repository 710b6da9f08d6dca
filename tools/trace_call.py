import os, sys
os.environ["HORAE_TRACE"] = "1"
sys.argv = [sys.argv[0], "none", "16", "5"]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profile_fused.py")).read())
