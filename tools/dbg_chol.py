import numpy as np, sys
sys.path.insert(0,'.')
import gamc_b200 as g, oracle as o
N,p=4096,256
X,y,tt=o.synth_logistic(1,N,p)
eng=g.Engine(0); eng.target_logistic(X,y,100.)
sampler=g.GAMC(update=g.smmala_only_update(),driftstep=0.02,minorscale=0.001,c=0.01,am_mode=1)
tuner=g.GAMCTuner(g.VanillaMCTuner(),g.VanillaMCTuner(),g.AcceptanceRateMCTuner(0.35))
job=g.BasicMCJob(g.LogisticModel(X,y,100.),sampler,g.BasicMCRange(1,0),{"p":tt},tuner=tuner,tape=g.RandomTape(1,p,1),engine=eng)
