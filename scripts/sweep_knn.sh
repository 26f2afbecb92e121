for cfg in "65536 12 knn 5" "4096 12 knn 5" "65536 8 knn 5" "65536 12 knn 10"; do for m in 0 1; do SWARM_TC=$m python scripts/sweep_rollout.py $cfg; done; done
