#!/usr/bin/env python
"""One-off pinning campaign at BASELINE.json's full sizes (build container, needs cv2; about a minute): every oracle stage against
cv2 4.13 on the 1920x1080 (seed 2) and 3840x2160 (seed 3) synthetic frames -- gray, medianBlur, Canny, dilate, the whole shape-method
chain, white->black + Laplacian sharpen, Otsu, distanceTransform (sharpened and raw masks), normalize, peaks, the whole
colour-method chain (contour labelling + circle), cv::watershed on those markers, pyrDown, pyrMeanShiftFiltering (1080p) and
bilateralFilter (max |diff|).  Result when run for round 1: every stage equal at both sizes, bilateral max |diff| = 1."""
import numpy as np, cv2, sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import oracle as orc
cv2.setNumThreads(8)
def canonical(l):
    flat=l.ravel(); pos=flat>0
    u,first=np.unique(flat[pos],return_index=True)
    order=np.argsort(first); rank=np.empty_like(order); rank[order]=np.arange(len(order))
    out=np.zeros_like(flat); out[pos]=rank[np.searchsorted(u,flat[pos])]+1
    return out.reshape(l.shape)
for (w,h,seed) in [(1920,1080,2),(3840,2160,3)]:
    t0=time.time()
    im=orc.synth_bgr(w,h,seed)
    res={}
    g=cv2.cvtColor(im,cv2.COLOR_BGR2GRAY); res['gray']=np.array_equal(g,orc.bgr2gray(im))
    k=orc.blur_mask_size(w,h)
    mb=cv2.medianBlur(g,k); res['median%d'%k]=np.array_equal(mb,orc.median_blur(g,k))
    res['median3']=np.array_equal(cv2.medianBlur(g,3),orc.median_blur(g,3))
    e=cv2.Canny(mb,5,50); res['canny']=np.array_equal(e,orc.canny(mb,5,50))
    d3=cv2.dilate(e,np.ones((3,3),np.uint8)); d5=cv2.dilate(d3,np.ones((5,5),np.uint8))
    res['dilate']=np.array_equal(d5,orc.dilate_rect(orc.dilate_rect(e,3,3),5,5))
    dde=cv2.medianBlur(cv2.subtract(d5,d3),3)
    n,l=cv2.connectedComponents(dde,connectivity=8,ltype=cv2.CV_32S)
    on,ol,_=orc.shape_seeds(im)
    res['shape_chain']=(n==on) and np.array_equal(canonical(l),ol)
    k91=np.array([1,1,1,1,-8,1,1,1,1],np.float32).reshape(9,1)
    black=im.copy(); black[(im==255).all(axis=2)]=0
    res['white_pixels']=int((im==255).all(axis=2).sum())
    sharp=np.clip(np.rint(black.astype(np.float32)-cv2.filter2D(black,cv2.CV_32F,k91)),0,255).astype(np.uint8)
    res['sharpen']=np.array_equal(sharp,orc.laplacian_sharpen(orc.white_to_black(im),orc.SHARPEN_TAPS_9x1))
    gs=cv2.cvtColor(sharp,cv2.COLOR_BGR2GRAY)
    t,bw=cv2.threshold(gs,40,255,cv2.THRESH_BINARY|cv2.THRESH_OTSU); res['otsu']=(int(t)==orc.otsu_threshold(gs))
    for name,mask in (('dt_sharp',bw),('dt_raw',cv2.threshold(g,0,255,cv2.THRESH_BINARY|cv2.THRESH_OTSU)[1])):
        res[name]=np.array_equal(cv2.distanceTransform(mask,cv2.DIST_L2,5),orc.distance_transform(mask))
    cn,cm,st=orc.color_seeds(im)
    dist=cv2.normalize(cv2.distanceTransform(bw,cv2.DIST_L2,5),None,0,1.,cv2.NORM_MINMAX)
    res['norm']=np.array_equal(dist,st['norm'])
    pk=cv2.dilate(cv2.threshold(dist,.4,1.,cv2.THRESH_BINARY)[1],np.ones((3,3),np.uint8)).astype(np.uint8)
    res['peaks']=np.array_equal(pk,st['peaks'])
    cs,hier=cv2.findContours(pk,cv2.RETR_CCOMP,cv2.CHAIN_APPROX_NONE)
    m=np.zeros(pk.shape,np.int32)
    for i in range(len(cs)): cv2.drawContours(m,cs,i,(i+1,)*4,-1,8,hier,2**31-1,(0,0))
    cv2.circle(m,(5,5),3,(255,255,255),-1)
    res['color_chain']=(len(cs)==cn) and np.array_equal(m,cm)
    # watershed with the colour-method markers (the reference's real region growing)
    mk=cm.copy(); cv2.watershed(sharp,mk)
    res['watershed']=np.array_equal(mk,orc.watershed(sharp,cm))
    # pyramid
    res['pyrDown']=np.array_equal(cv2.pyrDown(im),orc.pyr_down(im))
    if w==1920:
        f=cv2.pyrMeanShiftFiltering(im,10,10,maxLevel=1,termcrit=(3,5,1.0))
        res['meanshift']=np.array_equal(f,orc.meanshift_filter(im,10,10,1))
        # the floodFill region-growing loop (samples/cpp/meanshift_segmentation.cpp) on the filtered frame vs the oracle's labels
        fmask=np.zeros((h+2,w+2),np.uint8); flab=np.zeros((h,w),np.int32); nreg=0; inner=fmask[1:-1,1:-1]; fimg=f.copy()
        for yy in range(h):
            while True:
                xs=np.flatnonzero(inner[yy]==0)
                if len(xs)==0: break
                nreg+=1
                cv2.floodFill(fimg,fmask,(int(xs[0]),yy),(0,0,0),(2,2,2),(2,2,2),4|cv2.FLOODFILL_MASK_ONLY|(2<<8))
                new=(inner==2); flab[new]=nreg; inner[new]=1
        on2,ol2=orc.label_regions(f,2)
        res['floodfill_labels']=(nreg==on2) and np.array_equal(flab,ol2)
        b1=cv2.bilateralFilter(g,11,22,22); o1=orc.bilateral_filter(g,11,22,22)
        res['bilateral_maxdiff']=int(np.abs(b1.astype(int)-o1.astype(int)).max())
    print((w,h),'%.0f s'%(time.time()-t0),res)
