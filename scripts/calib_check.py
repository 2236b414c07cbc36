import ctypes, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "qwen-megakernel-tts_b200"))
from qwen_megakernel import build_tts
lib = build_tts.load_library()
h = ctypes.c_void_p()
t = time.time()
rc = lib.qmk_engine_create(0, 0, ctypes.byref(h))
print("engine_create rc", rc, "in", round(time.time() - t, 3), "s", lib.qmk_last_error().decode() if rc else "")
