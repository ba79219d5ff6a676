run base --args "--steps 20" A=1
run lab2 --args "--steps 20" RMCV_LABEL_MINSMEM=77000
run lab3 --args "--steps 20" RMCV_LABEL_MINSMEM=60000
run s2 --args "--steps 20" RMCV_PIX_S=2
run s2lab2 --args "--steps 20" RMCV_PIX_S=2 RMCV_LABEL_MINSMEM=77000
run s2lab3 --args "--steps 20" RMCV_PIX_S=2 RMCV_LABEL_MINSMEM=60000
run prio1 --args "--steps 20" RMCV_PRIO=1
run prio1s2 --args "--steps 20" RMCV_PRIO=1 RMCV_PIX_S=2
