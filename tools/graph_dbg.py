import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import openjph_b200 as ob
from openjph_b200 import _lib
L = _lib.lib()
W=H=512
p = ob.make_params(W,H,3,8,num_decomps=5,reversible=True,color_transform=True)
rng=np.random.default_rng(0)
fr=[rng.integers(0,256,(H,W)).astype(np.uint8) for _ in range(3)]
pin=[torch.from_numpy(f).pin_memory() for f in fr]
planes=(C.c_void_p*3)(*[t.data_ptr() for t in pin])
enc=L.ojb_enc_create(); dec=L.ojb_dec_create()
assert L.ojb_enc_configure(enc, C.byref(p), ob.U8)==0
assert L.ojb_enc_upload_frame(enc, planes, None)==0
cap=W*H*3*2+(1<<20)
cs=torch.empty(cap,dtype=torch.uint8,device="cuda"); n=C.c_uint64(); fi=_lib.FrameInfo()
for i in range(4):
    rc=L.ojb_enc_encode_resident(enc, cs.data_ptr(), cap, C.byref(n), 1)
    print("enc", i, rc, n.value, L.ojb_last_error() if rc else "")
    rc=L.ojb_dec_read_headers_device(dec, cs.data_ptr(), n.value, ob.U8, C.byref(fi)); print(" hdr", rc, L.ojb_last_error() if rc else "")
    rc=L.ojb_dec_decode_resident(dec); print(" dec", i, rc, L.ojb_last_error() if rc else "")
