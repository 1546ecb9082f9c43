"""numpy model of the merged Lanczos schedule of csrc/k4_chain.cu (lanczos_merged_kernel): the unorthogonalised candidate w_j is what the
ranks exchange, Ms q_{j+1} follows from y = Ms w_j by linearity and the three-term relation of the basis (no stored products).
Compared with the classical iteration (direct products) on the benchmark matrix (N = 2000, 250 x 250, SNR 0.1), a pristine N = 400
matrix and one with every image duplicated (rank-deficient): alpha / beta, drift |z - Ms q| / |Ms q|, orthogonality, Ritz residuals.
Test infrastructure (uses the oracle); takes about a minute.
    python tools/merged_lanczos_model.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
def run(Ms, nsteps, label, seed=0):
    m = Ms.shape[0]
    v0 = np.ones(m)
    for _ in range(300): v0 = Ms@v0; v0/=np.linalg.norm(v0)
    rng = np.random.default_rng(seed)
    q1 = rng.standard_normal(m); q1 -= v0*(v0@q1); q1/=np.linalg.norm(q1)
    def standard(nsteps):
        V = np.zeros((nsteps+2, m)); V[0]=v0; V[1]=q1
        al=[]; be=[]; sigma=0; beta_prev=0
        for j in range(1, nsteps+1):
            w = Ms@V[j] - sigma*V[j] - (beta_prev*V[j-1] if j>=2 else 0)
            h = V[:j+1]@w; ww = w@w
            w = w - V[:j+1].T@h
            alpha = sigma + h[j]; beta = np.sqrt(max(ww - h@h,0))
            al.append(alpha); be.append(beta)
            if beta < 1e-14*abs(al[0]): break
            V[j+1] = w/beta; sigma=alpha; beta_prev=beta
        return np.array(al), np.array(be), V
    def merged(nsteps):
        # T-based: Z_i = beta_i q_{i+1} + alpha_i q_i + beta_{i-1} q_{i-1}  (i>=1, beta_0 = 0 ... coupling to v0 none), Z_0 = v0 (lambda0 = 1)
        V = np.zeros((nsteps+2, m)); V[0]=v0; V[1]=q1
        al=[0.0]; be=[0.0]   # index by step: al[i], be[i] for i>=1
        drift=[]
        z1 = Ms@q1
        sigma=0.0
        w = z1 - sigma*V[1]
        for j in range(1, nsteps+1):
            h = V[:j+1]@w; ww = w@w
            y = Ms@w
            after = ww - h@h
            beta = np.sqrt(max(after,0)); alpha = sigma + h[j]
            al.append(alpha); be.append(beta)
            if beta < 1e-14*abs(al[1]): break
            q = (w - V[:j+1].T@h)/beta
            # g = coefficients of Z h in the basis V_0..V_j, plus beta_j h_j on q_{j+1}
            g = np.zeros(j+1)
            g[0] += 1.0*h[0]
            for i in range(1, j):       # i < j: three-term relation of earlier steps
                g[i] += al[i]*h[i]
                g[i+1] += be[i]*h[i]
                if i >= 2: g[i-1] += be[i-1]*h[i]
            # i = j: Z_j = beta_j q_{j+1} + sigma q_j + beta_{j-1} q_{j-1} + V h   (exact, this step)
            hj = h[j]
            g[j] += sigma*hj
            if j >= 2: g[j-1] += be[j-1]*hj
            g += h*hj
            z = (y - V[:j+1].T@g - beta*hj*q)/beta
            V[j+1]=q
            drift.append(np.linalg.norm(z - Ms@q)/np.linalg.norm(Ms@q))
            sigma=alpha
            w = z - sigma*q - beta*V[j]
        return np.array(al[1:]), np.array(be[1:]), V, np.array(drift)
    def ritz(al, be, V, k=10):
        n=len(al); T=np.diag(al)+np.diag(be[:n-1],1)+np.diag(be[:n-1],-1)
        th,S=np.linalg.eigh(T); idx=np.argsort(-th)[:k]
        th=th[idx]; S=S[:,idx]; U = V[1:n+1].T@S
        return th, U, np.linalg.norm(Ms@U - U*th, axis=0)
    a1,b1,V1 = standard(nsteps); a2,b2,V2,dr = merged(nsteps)
    n1=len(a1); n2=len(a2); n=min(n1,n2)
    t1,U1,r1 = ritz(a1,b1,V1); t2,U2,r2 = ritz(a2,b2,V2)
    print(label, 'steps', n1, n2, 'drift max %.2e last %.2e'%(dr.max(), dr[-1]),
          'ortho %.2e / %.2e'%(np.abs(V1[:n1+1]@V1[:n1+1].T-np.eye(n1+1)).max(), np.abs(V2[:n2+1]@V2[:n2+1].T-np.eye(n2+1)).max()))
    print('   res std %.2e merged %.2e; theta diff %.2e; alpha diff %.2e beta diff %.2e'%(r1.max(), r2.max(), np.abs(t1-t2).max()/abs(t1[0]), np.abs(a1[:n]-a2[:n]).max(), np.abs(b1[:n]-b2[:n]).max()))
from manifoldem_esper_b200 import synth
from oracle import esper_oracle as O
stack = synth.synthetic_pd(pd=1, box=250, n_side=20, tau=5, snr=0.1)
Ms = O.markov_symmetric_dense(O.distances_gram_f64(O.standardize_stack(stack)), 0.2834)
run(Ms, 200, 'C2'); run(Ms, 320, 'C2-long', 1)
st = synth.synthetic_pd(pd=1, box=64, n_side=20, tau=1, snr=None)
X = O.standardize_stack(st); D = O.distances_gram_f64(X)
Mp = O.markov_symmetric_dense(D, 0.0128)
run(Mp, 300, 'pristine400'); run(Mp, 399, 'pristine400-exh')
st2 = np.concatenate([st, st]); X = O.standardize_stack(st2); D = O.distances_gram_f64(X)
Md = O.markov_symmetric_dense(D, 0.0128)
run(Md, 420, 'dupes800')
