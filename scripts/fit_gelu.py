import numpy as np
from scipy.special import erfc, erf
# gelu(x) = relu(x) - |x| * 0.5*erfc(|x|/sqrt2) = relu(x) - |x| * exp2(q(|x|)),  q(t) = log2(0.5 erfc(t/sqrt2)), q(0) = -1
T=6.0
def fit(deg, iters=60):
    t=np.cos(np.linspace(0,np.pi,4001))*T/2+T/2
    y=np.log2(0.5*erfc(t/np.sqrt(2)))+1.0   # q(t)+1, zero at t=0
    wt=t*0.5*erfc(t/np.sqrt(2))*np.log(2)    # d gelu / d q
    wt=np.maximum(wt,1e-9)
    V=np.stack([t**k for k in range(1,deg+1)],1)
    w=wt.copy()
    for _ in range(iters):
        c,*_=np.linalg.lstsq(V*w[:,None],y*w,rcond=None)
        err=np.abs((V@c-y)*wt)
        w=w*(0.5+err/err.max())   # Lawson-ish
    return c, err.max()
for deg in (5,6,7,8,9):
    c,e=fit(deg)
    # evaluate in float32 Horner
    x=np.linspace(-8,8,400001).astype(np.float32)
    t=np.minimum(np.abs(x),np.float32(T))
    p=np.float32(c[-1])
    for k in range(deg-2,-1,-1):
        p=(p*t+np.float32(c[k])).astype(np.float32)
    p=(p*t-np.float32(1.0)).astype(np.float32)
    E=np.exp2(p.astype(np.float32)).astype(np.float32)
    g=(np.maximum(x,0)-np.abs(x)*E).astype(np.float32)
    ref=0.5*x.astype(np.float64)*(1+erf(x.astype(np.float64)/np.sqrt(2)))
    print(deg, "fit werr",e,"f32 max abs err",np.abs(g-ref).max(), "max rel (|ref|>1e-3)",(np.abs(g-ref)/np.maximum(np.abs(ref),1e-3)).max())
    if deg in (6,7,8): print("  coeffs c1..:", [float(np.float32(v)) for v in c])
import numpy as np
from scipy.special import erf
from scipy.optimize import least_squares
x=np.linspace(-7,7,56001)
ref=0.5*x*(1+erf(x/np.sqrt(2)))
def gelu_t(c,x,deg):
    x2=x*x
    p=c[-1]
    for k in range(len(c)-2,-1,-1): p=p*x2+c[k]
    u=x*p
    return 0.5*x*(1+np.tanh(u))
for deg in (2,3,4):
    c0=[0.7978845608,0.0356774081]+[0.0]*(deg-2)
    r=least_squares(lambda c: (gelu_t(c,x,deg)-ref)*1e3, c0[:deg], method='lm')
    # minimax refinement by reweighting
    c=r.x
    for it in range(30):
        e=np.abs(gelu_t(c,x,deg)-ref); w=(0.2+e/e.max())
        r=least_squares(lambda cc: (gelu_t(cc,x,deg)-ref)*w*1e3, c, method='lm'); c=r.x
    e=np.abs(gelu_t(c,x,deg)-ref)
    print(deg, 'max abs err',e.max(),'at x=',x[e.argmax()], 'coeffs',[float(np.float32(v)) for v in c])
# effect of tanh.approx 2^-11 relative error
c=r.x
