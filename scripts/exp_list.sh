run base --args "--steps 20" A=1
run c444 --args "--steps 20 --chunk 444" A=1
run rs2688 --args "--steps 20" RMCV_FRAME_RS=2688
run rs2560 --args "--steps 20" RMCV_FRAME_RS=2560
run c512 --args "--steps 20 --chunk 512" A=1
run c342 --args "--steps 20 --chunk 342" A=1
