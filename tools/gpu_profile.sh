#!/bin/bash
# Runs on the GPU box (via gpurun): ncu captures -> small CSV summaries in gpurun_out/.
# usage: tools/gpu_profile.sh <tag> [box] [launches] [conv]
set -u
export SSD_B200_PDL=0
export SSD_B200_IRBLOCK=1      # the block kernels of the production path (automatic mode keys on PDL, which is off under ncu)
tag=$1; shift
OUT=gpurun_out
mkdir -p $OUT
trap 'rm -f $OUT/*.ncu-rep' EXIT          # reports are far beyond what gpurun copies back: only the CSV summaries travel
METRICS='gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |dram__throughput.avg.pct_of_peak_sustained_elapsed|sm__throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|sm__pipe_tensor_cycles_active|sm__inst_executed_pipe_tensor|smsp__cycles_active.avg|l1tex__t_bytes|lts__t_bytes.sum |launch__occupancy_limit|smsp__average_warp.*_per_issue_active|sm__pipe_tensor_op_hmma_cycles_active'
for what in "$@"; do
  case $what in
    box)
      ncu --set full --clock-control none --import-source on -k regex:"iou_map|match_encode|loss_|nms_" \
          -o $OUT/box_$tag python tools/prof_box.py > $OUT/prof_box_$tag.log 2>&1
      ncu -i $OUT/box_$tag.ncu-rep --page raw --csv > $OUT/box_${tag}_raw.csv 2>/dev/null
      ncu -i $OUT/box_$tag.ncu-rep --page details --csv > $OUT/box_${tag}_details.csv 2>/dev/null
      python tools/ncu_summary.py $OUT/box_${tag}_raw.csv > $OUT/box_${tag}_summary.csv 2>&1
      ;;
    launches)
      ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_$tag.csv \
          python bench.py --profile-one-step > $OUT/launches_$tag.log 2>&1
      ;;
    conv)
      ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_tcgen05_kernel|conv_dwproj|conv_irblock|irblock_mma|stem_dwproj|conv_chain|stem_conv|conv_igemm|splitk|nms_|depthwise3x3" \
          -o $OUT/conv_$tag python bench.py --profile-one-step > $OUT/conv_$tag.log 2>&1
      ncu -i $OUT/conv_$tag.ncu-rep --page raw --csv > $OUT/conv_${tag}_raw.csv 2>/dev/null
      python tools/ncu_summary.py $OUT/conv_${tag}_raw.csv > $OUT/conv_${tag}_summary.csv 2>&1
      python tools/ncu_traffic.py $OUT/conv_${tag}_raw.csv $OUT/traffic_$tag.json > /dev/null 2>&1
      rm -f $OUT/conv_$tag.ncu-rep
      ;;
    vgg)
      # one SSD300-VGG16 B=32 inference step: tensor-pipe evidence for the tcgen05 / CTA-pair convolutions
      ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_tcgen05|stem_conv|conv_chain|maxpool|l2norm|splitk" \
          -o $OUT/vgg_$tag python tools/infer_bench.py --backbone vgg16 --batch 32 --profile-one-step > $OUT/vgg_$tag.log 2>&1
      ncu -i $OUT/vgg_$tag.ncu-rep --page raw --csv > $OUT/vgg_${tag}_raw.csv 2>/dev/null
      python tools/ncu_summary.py $OUT/vgg_${tag}_raw.csv > $OUT/vgg_${tag}_summary.csv 2>&1
      rm -f $OUT/vgg_$tag.ncu-rep $OUT/vgg_${tag}_raw.csv
      ;;
    train)
      # the training kernels of one MobileNetV2 step (BatchNorm fwd/bwd, depthwise gradients, tcgen05 wgrad, loss bwd, Adam);
      # one pass per kernel family so that each gets a sample from the middle of the network
      : > $OUT/train_${tag}_summary.csv
      for fam in "bn_:24:40" "dw_dgrad|dw_wgrad:16:8" "conv_wgrad:16:10" "adam_multi|loss_bwd|loss_anchor|loss_select:6:0"; do
        re=${fam%%:*}; rest=${fam#*:}; cnt=${rest%%:*}; skip=${rest#*:}
        ncu --set full --clock-control none --import-source on -k regex:"$re" -c $cnt --launch-skip $skip \
            -o $OUT/train_$tag python tools/train_bench.py --backbone mobilenet_v2 --breakdown > $OUT/train_$tag.log 2>&1
        ncu -i $OUT/train_$tag.ncu-rep --page raw --csv > $OUT/train_${tag}_raw.csv 2>/dev/null
        python tools/ncu_summary.py $OUT/train_${tag}_raw.csv >> $OUT/train_${tag}_summary.csv 2>&1
        rm -f $OUT/train_$tag.ncu-rep $OUT/train_${tag}_raw.csv
      done
      ;;
  esac
done
ls -la $OUT; du -sh $OUT
