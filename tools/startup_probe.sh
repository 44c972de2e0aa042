probe='import time;t=time.time();import torch;t1=time.time();torch.zeros(1,device="cuda:0");torch.cuda.synchronize();print("import %.2f cuda %.2f"%(t1-t,time.time()-t1))'
echo "--- default env, single process"; python -c "$probe"
echo "--- CUDA_VISIBLE_DEVICES=0"; CUDA_VISIBLE_DEVICES=0 python -c "$probe"
echo "--- OMP_NUM_THREADS=1"; OMP_NUM_THREADS=1 python -c "$probe"
echo "--- 2 concurrent, default"; (python -c "$probe" &) ; python -c "$probe"; sleep 8
echo "--- 2 concurrent, each own visible device"; (CUDA_VISIBLE_DEVICES=1 python -c "$probe" &) ; CUDA_VISIBLE_DEVICES=0 python -c "$probe"; sleep 8
echo "--- 2 concurrent, each own visible device, OMP 1"; (OMP_NUM_THREADS=1 CUDA_VISIBLE_DEVICES=1 python -c "$probe" &) ; OMP_NUM_THREADS=1 CUDA_VISIBLE_DEVICES=0 python -c "$probe"; sleep 8
nproc
