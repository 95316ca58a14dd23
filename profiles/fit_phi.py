import numpy as np
from numpy.polynomial import chebyshev as Ch
from math import erf, sqrt, pi
import scipy.special as sp
# fit Phi(x)-0.5 = x*P(x^2) on [0,R], minimize abs error; evaluate in float32 Horner
def fit(fun, R, deg, n=4000):
    # Chebyshev nodes in t=x^2 in [0,R^2]
    k = np.arange(n)
    t = 0.5*R*R*(1+np.cos(pi*(k+0.5)/n))
    x = np.sqrt(t)
    y = fun(x)/np.where(x==0,1,x)
    # least squares in Chebyshev basis on t, iterate weights for minimax-ish (weight by x to get abs error on x*P)
    A = Ch.chebvander(2*t/(R*R)-1, deg)
    w = x
    for it in range(30):
        c,*_ = np.linalg.lstsq(A*w[:,None], y*w, rcond=None)
        err = (A@c - y)*x
        # Lawson
        w = w*(1+ 0.5*np.abs(err)/np.abs(err).max())
    # convert to monomial in t
    p = Ch.cheb2poly(c)
    # p is in variable u=2t/R^2-1; convert to t
    from numpy.polynomial import polynomial as P
    u = np.array([-1, 2/(R*R)])
    mono = np.zeros(1)
    for i,ci in enumerate(p):
        mono = P.polyadd(mono, ci*P.polypow(u,i))
    return mono
def evalf32(mono, x):
    x = x.astype(np.float32); t = x*x
    acc = np.float32(mono[-1])*np.ones_like(x)
    for cI in mono[-2::-1]:
        acc = acc*t + np.float32(cI)
    return acc*x
Phi0 = lambda x: 0.5*sp.erf(x/np.sqrt(2))
xs = np.linspace(0,6,200001)
for R in (3.5,4.0,4.5):
  for deg in (5,6,7,8,9,10):
    m = fit(Phi0,R,deg)
    xc = np.minimum(xs,R)
    e = np.abs(evalf32(m,xc).astype(np.float64)-Phi0(xs)).max()
    print(R,deg,"maxabs err",e)
print("---- final")
R=4.5; deg=9
m = fit(Phi0,R,deg)
# scale so that f32 eval at R gives 0.5
v = float(evalf32(m, np.array([R]))[0]); m = m*(0.5/v)
xc = np.minimum(xs,R)
e = np.abs(evalf32(m,xc).astype(np.float64)-Phi0(xs))
print("max err", e.max(), "at", xs[e.argmax()], "val at R", evalf32(m,np.array([R]))[0])
print(", ".join("%.9ef"%c for c in m))
# gelu error rel to bf16
x = np.linspace(-8,8,400001)
xcl = np.clip(x,-R,R)
phi = 0.5+np.sign(xcl)*evalf32(m,np.abs(xcl)).astype(np.float64)
g = x*phi; gt = x*(0.5+Phi0(np.abs(x))*np.sign(x))
print("gelu max abs err", np.abs(g-gt).max(), "at", x[np.abs(g-gt).argmax()])
