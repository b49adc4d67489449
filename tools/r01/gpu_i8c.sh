#!/bin/bash
SPECS="full::12288 nofilter::12288:--dev-no-filter knnonly::12288:--dev-ratio,0.001 full_l::0 nofilter_l::0:--dev-no-filter knnonly_l::0:--dev-ratio,0.001 bp512::12288:--batch-pairs,512 bp128::12288:--batch-pairs,128" bash tools/gpu_i8ab.sh
