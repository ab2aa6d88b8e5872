#!/bin/bash
# On the B200 box: the reference's own train.py and inference.py (unchanged files from _ref_scratch/, staged by
# scripts/stage_reference.sh) through `gnn_bfs_rans_b200.dropin`, for every layer type, cfg1/cfg2 shape (hidden 128, L=4).
# Logs, checkpoints and predictions land in gpurun_out/refrun/.
cd "$(dirname "$0")/.."
OUT=$PWD/gpurun_out/refrun
mkdir -p $OUT
EPOCHS=${EPOCHS:-3}
for lt in GCN GAT GIN Transformer; do
  timeout 600 python scripts/ref_dropin_runner.py _ref_scratch/train.py --layer_type $lt --hidden_dim 128 --num_layers 4 \
      --epochs $EPOCHS --device cuda --output_dir $OUT/ckpt_$lt > $OUT/train_$lt.log 2>&1
  echo "train $lt rc=$?" | tee -a $OUT/status.txt
  timeout 600 python scripts/ref_dropin_runner.py _ref_scratch/inference.py --checkpoint $OUT/ckpt_$lt/best_model.pt \
      --device cuda --output_dir $OUT/pred_$lt > $OUT/infer_$lt.log 2>&1
  echo "inference $lt rc=$?" | tee -a $OUT/status.txt
  rm -f $OUT/ckpt_$lt/checkpoint_epoch_*.pt
done
grep -h B2G_DROPIN_SUMMARY $OUT/*.log > $OUT/summaries.txt
tail -n 3 $OUT/train_*.log | cut -c1-400
cat $OUT/status.txt
