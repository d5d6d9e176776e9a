set -e
mkdir -p /tmp/cli2 && cd /tmp/cli2 && rm -rf results output_log.txt
printf 'qc:4,5,10,61,9,49\n0\n0\n300000\n50\n0.05' > init.txt
$GRAFT_REPO_ROOT/qec_ldpc_b200/lib/qec_ldpc init.txt --depolarizing --seed 123 --gpus 2
$GRAFT_REPO_ROOT/qec_ldpc_b200/lib/qec_ldpc init.txt --depolarizing --seed 123 --gpus 1
cat results/*depolarizing*.txt | grep -E "Corrected|Logical|Syndrome|Tested:|Duration"
tail -3 output_log.txt
