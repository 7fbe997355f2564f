import sys, os; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
os.environ.setdefault('X','1')
import importlib.util
src = open('/root/repo/tools/time_conv.py').read().split("print('DBG'")[0]
exec(src)
print('DBG', os.environ.get('TGAN_IGEMM_DBG'))
run(250,32,128,128)
run(250,16,256,256)
