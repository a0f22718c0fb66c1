"""How many passes does a whole-set fixed-point iteration of greedy NMS need?  (CPU simulation behind nms_fixpoint_kernel, nms.cu.)
Plain Jacobi and the block variant the kernel runs (exact inside a chunk of 64, previous pass across chunks) on the bench's
6000 clustered boxes and on 6000 boxes jittered around 40 objects.  Measured: 6-10 passes, always the greedy answer."""
import numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskrcnn_b200 import synth
def iou_mat(b, thr):
    y1,x1,y2,x2 = b[:,0],b[:,1],b[:,2],b[:,3]
    area=(y2-y1+1)*(x2-x1+1)   # approximate; reference nms uses +1? doesn't matter for iteration counts
    yy1=np.maximum(y1[:,None],y1[None]); xx1=np.maximum(x1[:,None],x1[None])
    yy2=np.minimum(y2[:,None],y2[None]); xx2=np.minimum(x2[:,None],x2[None])
    inter=np.maximum(0,yy2-yy1+1)*np.maximum(0,xx2-xx1+1)
    return inter/(area[:,None]+area[None]-inter) >= thr
def greedy(M):
    n=len(M); keep=np.ones(n,bool)
    for i in range(n):
        if keep[i]: keep[i+1:] &= ~M[i,i+1:]
    return keep
def jacobi(M, chunk=None):
    n=len(M); L=np.tril(M,-1)   # L[i,j]=1 if j<i suppresses i
    keep=np.ones(n,bool); it=0
    while True:
        it+=1
        if chunk is None:
            new = ~(L.astype(np.int32)@keep.astype(np.int32) > 0)
        else:
            # block Gauss-Seidel within chunk using OLD values of earlier chunks, exact within chunk
            new=keep.copy()
            for c in range(0,n,chunk):
                e=min(n,c+chunk)
                sup = (L[c:e,:c].astype(np.int32)@keep[:c].astype(np.int32))>0
                k=~sup
                for i in range(c,e):
                    if k[i-c]:
                        for j in range(c,i):
                            if k[j-c] and L[i,j]: k[i-c]=False; break
                new[c:e]=k
        if (new==keep).all(): return keep,it
        keep=new
rng_=np.random.default_rng(11)
b=synth.random_rois(6000,11,image=1024.0,min_size=16,max_size=500)*1024.0
b[3000:]=b[:3000]+rng_.uniform(-8,8,(3000,4)).astype(np.float32)
M=iou_mat(b,0.7)
g=greedy(M); k,it=jacobi(M); print('bench jacobi iters',it,(k==g).all(),g.sum())
k,it=jacobi(M,64); print('bench blockGS iters',it,(k==g).all())
# dense clusters: 40 objects, 6000 boxes jittered
rng=np.random.default_rng(3)
for sig in (4,10,20,40):
    cen=synth.random_rois(40,5,image=1024.0,min_size=60,max_size=400)*1024.0
    bb=cen[rng.integers(0,40,6000)]+rng.normal(0,sig,(6000,4)).astype(np.float32)
    M=iou_mat(bb,0.7); g=greedy(M); k,it=jacobi(M); k2,it2=jacobi(M,64)
    print('sig',sig,'kept',g.sum(),'jacobi',it,(k==g).all(),'blockGS',it2,(k2==g).all())
