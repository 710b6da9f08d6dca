import os, sys
os.environ["HORAE_TRACE"] = "1"
sys.argv = [sys.argv[0], "snappy", "16", "4"]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profile_fused.py")).read())
