# Developer targets (reference: top-level Makefile with registry/build/push/deploy).
IMAGE ?= localhost:32000/adaptdl-b200-sched
TAG ?= dev
NAMESPACE ?= adaptdl
RELEASE ?= adaptdl-b200

.PHONY: build-native test test-gpu test-reference sanitize-host lint image push deploy delete config
build-native:
	python -c "import __graft_entry__ as g; g.build()"
test:
	python -m pytest tests -x -q -m "not gpu"
test-gpu:
	python -m pytest tests -x -q -m gpu
# the reference's own unit tests and example scripts, unmodified, on this framework
test-reference:
	python -m pytest tests/test_reference_suite.py tests/test_api_surface.py -q
# ASan + UBSan and TSan builds of csrc/host under the policy tests (no GPU needed)
sanitize-host:
	bash tools/sanitize_host.sh
lint:
	python -m flake8 adaptdl_b200 tests bench.py --max-line-length 100
image:
	docker build -f deploy/docker/Dockerfile.sched -t $(IMAGE):$(TAG) .
push: image
	docker push $(IMAGE):$(TAG)
deploy: push
	helm upgrade --install $(RELEASE) deploy/helm/adaptdl-b200-sched --namespace $(NAMESPACE) --create-namespace \
	    --set image.repository=$(IMAGE) --set image.tag=$(TAG)
delete:
	helm uninstall $(RELEASE) --namespace $(NAMESPACE)
config:
	kubectl get configmap $(RELEASE)-config --namespace $(NAMESPACE) -o yaml
