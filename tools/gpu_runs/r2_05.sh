set -x
for seg in 64 96 128 192 256; do BROADCAST_B200_MARCH_SEG=$seg python tools/res_one.py 8192x2048 0 10; done 2>&1 | grep variant
python tools/res_one.py 8192x2048 0 10
python tools/res_one.py 8192x2048 4 10
python tools/res_one.py 1024x2048 0 20
python tools/res_one.py 1024x2048 4 20
python tools/res_one.py 500x150 0 20
python tools/res_one.py 500x150 4 20
python tools/res_one.py 630x300 0 20
python tools/res_one.py 630x300 4 20
