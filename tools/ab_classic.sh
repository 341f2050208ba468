for cfg in c2x2 c2x3; do for i in 1 2; do for m in 0 103; do
  lib=dual-interactive-implicit-neural-network_b200/libdiinn_b200.so; [ $m != 0 ] && lib=dual-interactive-implicit-neural-network_b200/build/libdiinn_b200_abl$m.so
  echo -n "mask $m: "; DIINN_B200_LIB=$PWD/$lib timeout -k 10 100 python tools/run_decode.py $cfg fp16 50 | tail -1
done; done; done
for i in 1 2; do for m in 0 103; do
  lib=dual-interactive-implicit-neural-network_b200/libdiinn_b200.so; [ $m != 0 ] && lib=dual-interactive-implicit-neural-network_b200/build/libdiinn_b200_abl$m.so
  echo -n "mask $m: "; DIINN_B200_LIB=$PWD/$lib timeout -k 10 100 python tools/run_decode.py c3 fp16 20 3 1 | tail -1
  echo -n "mask $m: "; DIINN_B200_LIB=$PWD/$lib timeout -k 10 100 python tools/run_decode.py c3 fp16 20 1 | tail -1
done; done
