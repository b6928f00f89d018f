#!/bin/sh
# Run ON THE GPU BOX (through gpurun): full-size vgg16 (2^25-entry input layer, synthetic weights) proved by the compiled reference
# (oracle/_ref/ref_run, ~70 s of CPU) and by zkcnn_prove on the same input and seed; the two transcripts must be byte-identical.
mkdir -p gpurun_out
CFG="64 64 M 128 128 M 256 256 256 M 512 512 512 M 512 512 512 M"
python tools/gen_synthetic_input.py vgg16 /tmp/vgg16.csv > /dev/null 2>&1
ls -la /tmp/vgg16.csv* | head
( timeout 420 oracle/_ref/ref_run vgg /tmp/vgg16.csv x /tmp/vgg16.csv.config 1 5 --transcript /tmp/ref16.bin > /tmp/ref16.log 2> /tmp/ref16.err ) &
timeout 600 zkcnn_b200/lib/zkcnn_prove vgg /tmp/vgg16.csv "$CFG" 1 5 --transcript /tmp/our16.bin --repeat 3 > /tmp/our16.log 2>&1
tail -3 /tmp/our16.log
wait
grep RESULT /tmp/ref16.log | cut -c1-200; tail -2 /tmp/ref16.err; tail -2 /tmp/ref16.log | cut -c1-200
if cmp -s /tmp/ref16.bin /tmp/our16.bin; then echo "VGG16 TRANSCRIPTS IDENTICAL ($(stat -c %s /tmp/our16.bin) bytes)"; else echo "VGG16 TRANSCRIPTS DIFFER"; fi
( echo "vgg16 (synthetic weights, pic_cnt = 1, seed 5, degenerate generators): reference CPU prover vs zkcnn_b200 on the same GPU box"; grep RESULT /tmp/ref16.log | cut -c1-220; tail -3 /tmp/our16.log; cmp /tmp/ref16.bin /tmp/our16.bin && echo "transcripts identical: $(stat -c %s /tmp/our16.bin) bytes, sha256 $(sha256sum /tmp/our16.bin | cut -c1-16)" ) > gpurun_out/r01_vgg16_parity.txt 2>&1
